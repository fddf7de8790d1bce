import sys, torch, math
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import c2dsr_oracle as oracle
from c2dsr_b200 import ops
from test_gpu_kernels import _seq_with_pads, _weight_list, _encoder_weights
from helpers import rel_err
DEV='cuda'
B, L, d, H, nl, nf = 48, 15, 256, 1, 1, False
g = torch.Generator().manual_seed(B * L + d + H)
pad = 999
seq = _seq_with_pads(B, L, 900, pad, g, full_rows=2)
x = torch.randn(B, L, d, generator=g); go = torch.randn(B, L, d, generator=g)
W = _encoder_weights(d, nl, g)
Wc = {k: v.clone().double().requires_grad_(True) for k, v in W.items()}
xc = x.clone().double().requires_grad_(True)
ref = oracle.encoder(xc, seq, Wc, "enc", pad, H, nl, nf)
ref.backward(go.double())
names = ["in_w","in_b","out_w","out_b","l1_w","l1_b","l2_w","l2_b","n1_w","n1_b","n2_w","n2_b","nf_w","nf_b"]
for dense in (0, 3, 1):
    wl = [t.to(DEV).requires_grad_(True) for t in _weight_list(W, nl)]
    xg = x.to(DEV).requires_grad_(True)
    out = ops.EncoderFn.apply(xg, seq.to(DEV), H, pad, nf, 0.0, 0, 0, dense, *wl)
    out.backward(go.to(DEV))
    print('dense', dense, 'out', rel_err(out.detach().cpu(), ref.detach()), 'dx', rel_err(xg.grad.cpu(), xc.grad))
    print('   ', {n: '%.1e' % rel_err(w.grad.cpu(), r.grad) for n, w, r in zip(names, wl, _weight_list(Wc, nl))})
# direct GEMM checks at encoder shapes
T = B * L
for (ta, tb, M, N, K) in [(0,1,T,768,256),(0,0,T,256,768),(1,0,768,256,T),(0,0,T,256,256),(1,0,256,256,T)]:
    A = torch.randn((K, M) if ta else (M, K), generator=g); Bm = torch.randn((N, K) if tb else (K, N), generator=g)
    refm = (A.t() if ta else A).double() @ (Bm.t() if tb else Bm).double()
    C = torch.zeros(M, N, device=DEV)
    ops.gemm_tc(ta, tb, M, N, K, A.to(DEV), A.shape[1], Bm.to(DEV), Bm.shape[1], C, N)
    print('gemm', ta, tb, M, N, K, rel_err(C.cpu(), refm))
