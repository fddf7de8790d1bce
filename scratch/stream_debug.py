"""Compare branch_streams on/off: losses and grads at a sweep corner."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import helpers  # noqa
import torch
import test_gpu_configs as T
from c2dsr_b200 import ops
V = os.environ.get("V", "")
_gb, _eb = ops.GatherFn.backward, ops.EncoderFn.backward
def gb(ctx, dx):
    cur, dflt = torch.cuda.current_stream(), torch.cuda.default_stream()
    if "1" in V: cur.wait_stream(dflt)
    out = _gb(ctx, dx)
    if "2" in V: dflt.wait_stream(cur)
    if "4" in V:
        for t in out:
            if t is not None: t.record_stream(dflt)
    return out
def eb(ctx, d):
    cur, dflt = torch.cuda.current_stream(), torch.cuda.default_stream()
    if "3" in V: cur.wait_stream(dflt)
    if "5" in V: d.record_stream(cur)
    out = _eb(ctx, d)
    if "6" in V: dflt.wait_stream(cur)
    return out
ops.GatherFn.backward = staticmethod(gb)
ops.EncoderFn.backward = staticmethod(eb)

for shape in [(512, 15, 1, 1, 32)]:
    res = {}
    for bs in (False, True, True, True):
        tr, otr, batch, ebatch = T._setup(*shape)
        tr.model.branch_streams = bs
        tr.model.train(); tr.optimizer.zero_grad()
        out = []
        for step in range(2):
            tr.model.convolve_graph()
            l = [float(x) for x in tr.train_batch(batch)]
            g = {k: v.grad.detach().clone() for k, v in tr.model.named_parameters() if v.grad is not None}
            out.append((l, g))
        torch.cuda.synchronize()
        if not bs:
            res = out
        else:
            for step in range(2):
                worst = sorted(((float((out[step][1][k] - res[step][1][k]).abs().max() /
                                (res[step][1][k].abs().max() + 1e-30)), k) for k in res[step][1]), reverse=True)[:5]
                print("   worst grads", worst)
