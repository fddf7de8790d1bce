"""Build-container check: the oracle port's CPU step time next to the reference's own code.

    python oracle/compare_port_timing.py        # needs /root/reference; CPU only; ~2 min

bench.py reports `cpu_baseline.kind = "port"` because the reference is Python source that cannot
travel to the GPU box.  This script times, on the same host and the same Food-Kitchen-shaped batch,
(a) the reference's Trainer.train_batch + convolve_graph and evaluate_batch and (b) the oracle port,
so that the port can be read as a stand-in for the reference's CPU path.
"""
import argparse, os, sys, time, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench, c2dsr_oracle as oracle

hp = bench.hyper(bench.WORKLOADS["fk"], 0.2, torch.device("cpu"))
adj, fields, ev = bench.make_workload(hp, 3, 1, seed=0)
B = hp.batch_size
batches = [tuple(torch.from_numpy(np.ascontiguousarray(fields[i * B:(i + 1) * B, f])) for f in range(14)) for i in range(3)]
six, four, neg = ev
nq = 256
eb = tuple(torch.from_numpy(np.ascontiguousarray(six[:nq, f])) for f in range(6)) + \
    tuple(torch.from_numpy(np.ascontiguousarray(four[:nq, f:f + 1])) for f in range(4)) + (torch.from_numpy(np.ascontiguousarray(neg[:nq])),)

# ---- (a) the reference itself -----------------------------------------------------------------
sys.path.insert(0, os.environ.get("C2DSR_REFERENCE", "/root/reference"))
from models.C2DSR import C2DSR
from trainer import Trainer
args = argparse.Namespace(**{k: v for k, v in vars(hp).items()})
args.device = torch.device("cpu")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    tr = Trainer.__new__(Trainer)
    tr.model = C2DSR(args, adj[0].coalesce(), adj[1].coalesce())
    tr.optimizer = torch.optim.AdamW(tr.model.parameters(), lr=args.lr, weight_decay=args.l2, amsgrad=True)
    tr.device, tr.d_latent, tr.n_item_a, tr.n_item_b = args.device, args.d_latent, args.n_item_a, args.n_item_b
    tr.len_rec, tr.lambda_loss = args.len_rec, args.lambda_loss
    tr.label_pos, tr.label_neg = torch.ones(B, 1), torch.zeros(B, 1)
    tr.model.train(); tr.optimizer.zero_grad()
    def ref_step(b):
        tr.model.convolve_graph(); return tr.train_batch(b)
    ref_step(batches[0])
    t0 = time.perf_counter(); ref_step(batches[1]); ref_step(batches[2]); t_ref = (time.perf_counter() - t0) / 2
    tr.model.eval()
    with torch.no_grad():
        tr.model.convolve_graph()
        t0 = time.perf_counter(); tr.evaluate_batch(eb); t_ref_ev = time.perf_counter() - t0

# ---- (b) the oracle port ---------------------------------------------------------------------------
h = {k: v for k, v in vars(hp).items() if isinstance(v, (int, float, bool, str))}
otr = oracle.OracleTrainer(oracle.init_state(h, seed=1), adj[0].coalesce(), adj[1].coalesce(), h)
otr.zero_grad(); otr.train_batch(batches[0], training=True)
t0 = time.perf_counter(); otr.train_batch(batches[1], training=True); otr.train_batch(batches[2], training=True)
t_port = (time.perf_counter() - t0) / 2
otr.convolve_graph()
t0 = time.perf_counter(); otr.evaluate_batch(eb); t_port_ev = time.perf_counter() - t0
print(f"threads {torch.get_num_threads()}  train step: reference {t_ref:.2f} s ({B / t_ref:.1f} seq/s)  port {t_port:.2f} s "
      f"({B / t_port:.1f} seq/s)  ratio port/ref {t_port / t_ref:.2f}")
print(f"eval {nq} queries (999 negatives): reference {t_ref_ev:.2f} s ({nq / t_ref_ev:.0f} q/s)  port {t_port_ev:.2f} s "
      f"({nq / t_port_ev:.0f} q/s)")
