"""Kineto timeline of full-catalogue evaluation batches (FK shape, one GPU)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
from c2dsr_b200.trainer import Trainer
import numpy as np
dev = torch.device("cuda", 0)
sys.argv = ["bench.py"]
hp = bench.hyper(bench.WORKLOADS[os.environ.get("WL", "fk")], 0.2, dev)
adj, fields, ev = bench.make_workload(hp, 1, 4, seed=0)
ds = CDSRDataset.from_fields([fields[:, i] for i in range(14)], "train", hp.len_max)
tr = Trainer.from_parts(hp, bench.Quiet(), (BatchLoader(ds, hp.batch_size), None, None), adj[0], adj[1])
ebs = bench.eval_batches(ev, hp.batch_size_eval, True)
dev_eb = [tuple(x.to(dev) for x in b[:10]) + (b[10],) for b in ebs]
tr.enable_pad_shortcut(ebs)
tr.model.eval()
with torch.no_grad():
    tr.model.convolve_graph()
    for i in range(4): tr.evaluate_batch(dev_eb[i % 4])
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    N = 8
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in tr.evaluate_stream(dev_eb[i % 4] for i in range(N)): pass
        torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t_first = evs[0].time_range.start
span = (evs[-1].time_range.end - t_first) / N
print("span us/batch", span, "busy us/batch", sum(e.device_time for e in evs) / N)
last = [e for e in evs if e.time_range.start >= t_first + span * (N - 2) - 2 and e.time_range.start < t_first + span * (N - 1)]
t0 = last[0].time_range.start
for e in last:
    print(f"S{getattr(e, 'device_resource_id', 0):<4d} {e.time_range.start - t0:8.1f} {e.device_time:7.1f}  {e.name[:90]}")
