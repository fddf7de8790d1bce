"""Stand-alone timing of the CSR SpMM on the FK-shaped synthetic graph."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from c2dsr_b200 import ops
from c2dsr_b200.graph import CsrGraph
dev = torch.device("cuda", 0)
hp = bench.hyper(bench.WORKLOADS["fk"], 0.2, dev)
adj, fields, ev = bench.make_workload(hp, 2, 1, seed=0)
g = CsrGraph(adj[0], dev)
n, d = hp.n_item, hp.d_latent
X = torch.randn(n, d, device=dev); Y = torch.randn(n, d, device=dev); out = torch.empty_like(X)
nnz = g.fwd[1].numel()
alg = nnz * (8 + 4 * d) + (n + 1) * 4 + 2 * n * d * 4
def t(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
cases = {
  "fwd p=0 +Y": lambda: ops.spmm(g.fwd, X, Y=Y, out=out, alpha=0.5, beta=0.5),
  "fwd p=0 noY": lambda: ops.spmm(g.fwd, X, out=out),
  "fwd p=.2 mode1 +Y": lambda: ops.spmm(g.fwd, X, Y=Y, out=out, alpha=0.5, beta=0.5, drop_mode=1, p=0.2, seed=1, tag=1),
  "bwd p=.2 mode2 +Y": lambda: ops.spmm(g.bwd, X, Y=Y, out=out, alpha=0.5, beta=0.5, drop_mode=2, p=0.2, seed=1, tag=1),
  "copy X->out (torch)": lambda: out.copy_(X),
}
print("n", n, "nnz", nnz, "alg MB", alg / 1e6, "kernel", os.environ.get("C2DSR_SPMM", "default"))
for k, f in cases.items():
    us = t(f)
    print(f"{k:24s} {us:8.1f} us   {alg / us / 1e3:8.1f} GB/s algorithmic")
