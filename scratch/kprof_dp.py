"""Kineto timeline of one data-parallel training step (rank 0 prints): torchrun --nproc-per-node N scratch/kprof_dp.py"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from c2dsr_b200 import dist as cdist
from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
from c2dsr_b200.trainer import Trainer
rank, world, local_rank = cdist.init_from_env("nccl")
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
sys.argv = ["bench.py"]
hp = bench.hyper(bench.WORKLOADS["fk"], 0.2, dev)
adj, fields, ev = bench.make_workload(hp, 8 * world, 1, seed=0)
fields = fields[rank::world]
ds = CDSRDataset.from_fields([fields[:, i] for i in range(14)], "train", hp.len_max)
loader = BatchLoader(ds, hp.batch_size, len_rec=hp.len_rec, ignore=(hp.n_item_a, hp.n_item_b))
tr = Trainer.from_parts(hp, bench.Quiet(), (loader, None, None), adj[0], adj[1])
ds.to(dev)
tb = list(loader)
for b in tb:
    b.global_rows = b.global_batch = hp.batch_size * world
tr.model.train(); tr.optimizer.zero_grad()
for i in range(4): tr.train_step(tb[i % len(tb)])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N): tr.train_step(tb[i % len(tb)])
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    t_first = evs[0].time_range.start
    span = (evs[-1].time_range.end - t_first) / N
    print("span us/step", span)
    last = [e for e in evs if e.time_range.start >= t_first + span * (N - 1) - 5]
    t0 = last[0].time_range.start
    for e in last:
        if e.device_time >= 15 or "nccl" in e.name.lower() or "barrier" in e.name.lower() or "peer" in e.name.lower():
            print(f"S{getattr(e, 'device_resource_id', 0):<4d} {e.time_range.start - t0:8.1f} {e.device_time:7.1f}  {e.name[:80]}")
import gc
tr._graphs.clear(); gc.collect(); torch.cuda.synchronize()
sys.stdout.flush(); os._exit(0)
