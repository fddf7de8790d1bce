"""One training step + one eval batch between cudaProfilerStart/Stop (for ncu --profile-from-start off) of a training step at the FK bench shape."""
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
tr.use_graph = False
tr.model.train(); tr.optimizer.zero_grad()
def step(i):
    tr.model.convolve_graph()
    tr.train_batch(tb[i % len(tb)])
ebs = bench.eval_batches(ev, hp.batch_size_eval, False)
eb = tuple(x.to(dev) for x in ebs[0])
def ev_step():
    tr.model.eval()
    with torch.no_grad():
        tr.model.convolve_graph()
        tr.evaluate_batch(eb)
    tr.model.train()
for i in range(3): step(i)
ev_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(3)
ev_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
