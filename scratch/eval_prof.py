import sys, time, torch, json
sys.path.insert(0, '.')
import bench
from c2dsr_b200 import _cabi
from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
from c2dsr_b200.trainer import Trainer
dev = torch.device('cuda', 0)
wl = bench.WORKLOADS['fk']; hp = bench.hyper(wl, 0.2, dev)
adj, fields, ev = bench.make_workload(hp, 2, 2, seed=0)
ds = CDSRDataset.from_fields([fields[:, i] for i in range(14)], "train", hp.len_max)
tr = Trainer.from_parts(hp, bench.Quiet(), (BatchLoader(ds, hp.batch_size), None, None), adj[0], adj[1])
eb = [tuple(x.to(dev) for x in b) for b in bench.eval_batches(ev, hp.batch_size_eval, True)]
tr.model.eval()
with torch.no_grad():
    tr.model.convolve_graph()
    for i in range(3): tr.evaluate_batch(eb[i % 2])
    torch.cuda.synchronize()
    _cabi.PROFILE = {"names": None, "events": []}
    t0 = time.perf_counter()
    for i in range(4): tr.evaluate_batch(eb[i % 2])
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 4
    agg = {}
    for nm, e0, e1 in _cabi.PROFILE["events"]:
        agg[nm] = agg.get(nm, 0.0) + e0.elapsed_time(e1) / 4
    _cabi.PROFILE = None
print('wall ms/batch', wall * 1e3, 'sum of kernels', sum(agg.values()))
print(json.dumps({k: round(v, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}))
from torch.profiler import profile, ProfilerActivity
with torch.no_grad(), profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2): tr.evaluate_batch(eb[i % 2])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
