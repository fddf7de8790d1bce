"""Attention fwd/bwd consistency under dropout: the tiled kernels vs the per-query fallback (forced by a
misaligned qkv view), and a double-precision finite difference of the forward."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from c2dsr_b200._cabi import call, ptr, stream
dev = "cuda"
torch.manual_seed(0)
B, L, d, H, pad, p = 16, 10, 64, 2, 999, 0.3
seq = torch.randint(0, 900, (B, L), device=dev)
for b in range(B):
    seq[b, : 1 + b % (L - 1)] = pad
def run(qkv, d_o, aligned=True):
    if aligned:
        q = qkv.contiguous()
        o = torch.empty(B, L, d, device=dev); dq = torch.empty_like(q)
    else:                                   # shift every buffer by one float: the tiled path refuses, fallback runs
        buf = torch.empty(qkv.numel() + 1, device=dev); q = buf[1:].view_as(qkv); q.copy_(qkv)
        ob = torch.empty(B * L * d + 1, device=dev); o = ob[1:].view(B, L, d)
        db = torch.empty(qkv.numel() + 1, device=dev); dq = db[1:].view_as(qkv)
    lse = torch.empty(B, L, H, device=dev)
    call("c2dsr_attention_fwd", q.data_ptr(), ptr(seq), B, L, d, H, pad, p, 77, 5, o.data_ptr(), ptr(lse), stream())
    call("c2dsr_attention_bwd", q.data_ptr(), o.data_ptr(), ptr(lse), ptr(d_o), ptr(seq), B, L, d, H, pad, p, 77, 5,
         dq.data_ptr(), stream())
    return o.clone(), dq.clone()
qkv = torch.randn(B, L, 3 * d, device=dev)
d_o = torch.randn(B, L, d, device=dev)
o1, g1 = run(qkv, d_o, True)
o2, g2 = run(qkv, d_o, False)
print("fwd tiled vs fallback", float((o1 - o2).abs().max()), "bwd", float((g1 - g2).abs().max()), "scale", float(g2.abs().max()))
v = torch.randn_like(qkv)
for eps in (1e-2, 1e-3):
    fp, _ = run(qkv + eps * v, d_o); fm, _ = run(qkv - eps * v, d_o)
    num = float(((fp.double() - fm.double()) * d_o.double()).sum() / (2 * eps))
    print("eps", eps, "numeric", num, "analytic", float((g1.double() * v.double()).sum()))
