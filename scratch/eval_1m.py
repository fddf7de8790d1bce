"""BASELINE configs[3]: synthetic 1M-item cross-domain catalogue, full-itemset evaluation (eval only), catalogue
sharded across the ranks of the job.  Prints one JSON line (rank 0).  Run with python (1 GPU) or torchrun."""
import sys, os, json, time, argparse
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
import bench
from c2dsr_b200 import _cabi, synth, dist as cdist
from c2dsr_b200.trainer import Trainer
from c2dsr_b200.dataloader import preprocess_evaluate
from c2dsr_b200.graph import normalised_coo, transition_edges, _to_sparse

rank, world, local_rank = cdist.init_from_env("nccl")
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
bench.WORKLOADS["1m"] = dict(shape="1m", d_latent=256, batch_size=256, batch_size_eval=2048)
hp = bench.hyper(bench.WORKLOADS["1m"], 0.0, dev)
sh = synth.SHAPES["1m"]
n_batches = int(os.environ.get("EVAL_BATCHES", "4"))
seqs = synth.make_sequences(200_000, hp.n_item_a, hp.n_item_b, len_max=hp.len_max, frac_a=sh["frac_a"], seed=0, lengths="fk")
shared, specific = transition_edges(seqs, hp.n_item_a)
adj = (_to_sparse(normalised_coo(shared, hp.n_item), hp.n_item), _to_sparse(normalised_coo(specific, hp.n_item), hp.n_item))
ev = synth.make_sequences(n_batches * hp.batch_size_eval, hp.n_item_a, hp.n_item_b, len_max=hp.len_max, frac_a=sh["frac_a"],
                          seed=1, lengths="fk")
import random; random.seed(0)
host_eb = bench.eval_batches(preprocess_evaluate(ev, hp.n_item_a, hp.n_item_b, hp.len_max, hp.n_neg_sample), hp.batch_size_eval, True)
dev_eb = [tuple(x.to(dev) for x in b) for b in host_eb]
torch.manual_seed(hp.seed)
tr = Trainer.from_parts(hp, bench.Quiet(), (None, None, None), adj[0], adj[1])
tr.model.eval()
with torch.no_grad():
    tr.model.convolve_graph()
    f = lambda bs: (lambda i: tr.evaluate_batch(bs[i % len(bs)]))
    for i in range(2): f(dev_eb)(i)
    names = {"c2dsr_score_count_tc", "c2dsr_score_target_tc"}
    _cabi.PROFILE = {"names": names, "events": []}
    n = 2 * n_batches
    ms = bench.timed(f(dev_eb), n, world)
    prof, _cabi.PROFILE = _cabi.PROFILE, None
    ms_h = bench.timed(f(host_eb), n, world)
k_ms = sum(a.elapsed_time(b) for _, a, b in prof["events"]) / n
flops = 2.0 * hp.d_latent * hp.batch_size_eval * (0.5 * hp.n_item_a + 0.5 * hp.n_item_b) / world
pk = bench.peaks()
if rank == 0:
    print(json.dumps({"metric": "full_catalog_eval_queries_per_sec", "value": round(n * hp.batch_size_eval / (ms / 1e3), 1),
                      "unit": "queries/s", "n_gpus": world, "ms_per_batch": round(ms / n, 3),
                      "config": {"workload": "synthetic 1M-item catalogue (400k + 600k items), d=256, L=15, 2048 queries per batch, "
                                             "catalogue rows sharded across ranks", "n_item": hp.n_item},
                      "e2e": {"value": round(n * hp.batch_size_eval / (ms_h / 1e3), 1), "unit": "queries/s"},
                      "roofline": {"kernel": "K4b score + rank count (per rank shard)", "bound": "tensor",
                                   "achieved": round(flops / (k_ms / 1e3) / 1e12, 2), "peak": pk["tensor_burst"] if "tensor_burst" in pk else pk["tensor"],
                                   "unit": "TFLOP/s", "ms_per_batch": round(k_ms, 3), "note": "algorithmic FLOPs of this rank's shard; bf16x3 executes 3x"}}), flush=True)
if world > 1:
    torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
