import sys, torch, math
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import c2dsr_oracle as oracle
from c2dsr_b200 import ops
from test_gpu_kernels import _seq_with_pads, _weight_list, _encoder_weights
DEV='cuda'
g = torch.Generator().manual_seed(3)
B, L, d, H, nl, pad = 16, 10, 64, 2, 2, 999
seq = _seq_with_pads(B, L, 900, pad, g)
x = torch.randn(B, L, d, generator=g)
W = _encoder_weights(d, nl, g)
c = torch.randn(B, L, d, generator=g); v = torch.randn(B, L, d, generator=g)
# oracle double
Wd = {k: t.double() for k, t in W.items()}
xd = x.double().requires_grad_(True)
fo = lambda xx: (oracle.encoder(xx, seq, Wd, "enc", pad, H, nl, False) * c.double()).sum()
fo(xd).backward()
an_o = float((xd.grad * v.double()).sum())
for eps in (1e-2, 4e-3, 1e-3, 1e-4):
    num_o = float((fo(x.double() + eps * v.double()) - fo(x.double() - eps * v.double())) / (2 * eps))
    print('oracle64 eps', eps, 'analytic', an_o, 'numeric', num_o)
wl = [t.to(DEV) for t in _weight_list(W, nl)]
for p in (0.0, 0.3):
    xg = x.to(DEV).requires_grad_(True)
    out = ops.EncoderFn.apply(xg, seq.to(DEV), H, pad, False, p, 1234, 5, 3, *wl)
    (out * c.to(DEV)).sum().backward()
    an = float((xg.grad.double() * v.to(DEV).double()).sum())
    f = lambda xx: float((ops.EncoderFn.apply(xx, seq.to(DEV), H, pad, False, p, 1234, 5, *wl).double() * c.to(DEV).double()).sum())
    for eps in (1e-2, 4e-3, 1e-3):
        num = (f(x.to(DEV) + eps * v.to(DEV)) - f(x.to(DEV) - eps * v.to(DEV))) / (2 * eps)
        print('gpu p', p, 'eps', eps, 'analytic', an, 'numeric', num)
