"""Kernel-level device time (kineto) of a training step at the FK bench shape."""
import sys, os, cProfile, pstats, argparse
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
sys.argv = ["bench.py"]
a = bench.parse()
from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
from c2dsr_b200.trainer import Trainer
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["fk"]
hp = bench.hyper(wl, 0.2, dev)
hp.score_path, hp.tc_passes = "tc", 3
adj, fields, ev = bench.make_workload(hp, 8, 1, seed=0)
ds = CDSRDataset.from_fields([fields[:, i] for i in range(14)], "train", hp.len_max)
loader = BatchLoader(ds, hp.batch_size, len_rec=hp.len_rec, ignore=(hp.n_item_a, hp.n_item_b))
tr = Trainer.from_parts(hp, bench.Quiet(), (loader, None, None), adj[0], adj[1])
ds.to(dev)
tb = list(loader)
tr.model.train(); tr.optimizer.zero_grad()
def step(i):
    tr.train_step(tb[i % len(tb)])
for i in range(3): step(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(N): step(i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    k = e.name[:90]
    a_ = agg.setdefault(k, [0, 0.0]); a_[0] += 1; a_[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print("total device-busy us/step (sum over streams):", tot / N)
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
print("span us/step:", (t1 - t0) / N)
by_stream = {}
for e in ev:
    by_stream.setdefault(getattr(e, "device_resource_id", getattr(e, "device_index", 0)), [0, 0.0])
    by_stream[getattr(e, "device_resource_id", getattr(e, "device_index", 0))][0] += 1
    by_stream[getattr(e, "device_resource_id", getattr(e, "device_index", 0))][1] += e.device_time
# ordered kernel sequence of the last step on the main stream
main = max(by_stream, key=lambda k: by_stream[k][1])
evs = sorted([e for e in ev if getattr(e, "device_resource_id", 0) == main], key=lambda e: e.time_range.start)
evs = evs[-(len(evs) // N):]
prev = None
print("---- main-stream sequence (start us rel, dur us, gap us, name)")
t00 = evs[0].time_range.start
for e in evs:
    gap = (e.time_range.start - prev) if prev is not None else 0
    print(f"{e.time_range.start - t00:8.1f} {e.device_time:7.1f} {gap:6.1f}  {e.name[:70]}")
    prev = e.time_range.end
print("---- all streams, last step, first 520 us (stream, start, dur, name)")
allv = sorted([e for e in ev if e.time_range.start >= t00], key=lambda e: e.time_range.start)
for e in allv:
    if e.time_range.start - t00 > 520: break
    print(f"S{getattr(e, 'device_resource_id', 0):<4d} {e.time_range.start - t00:8.1f} {e.device_time:7.1f}  {e.name[:60]}")
print("---- all streams, last step, kernels >= 12 us (stream, start, dur, name)")
for e in allv:
    if e.device_time >= 12:
        print(f"S{getattr(e, 'device_resource_id', 0):<4d} {e.time_range.start - t00:8.1f} {e.device_time:7.1f}  {e.name[:70]}")
print("per-stream busy us/step:", {k: (v[0] / N, round(v[1] / N, 1)) for k, v in by_stream.items()})
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{v[1]/N:9.1f} us  x{v[0]/N:5.1f}  {k}")
