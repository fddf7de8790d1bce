"""Kernel-level parity on the B200: every C-ABI entry point against the CPU oracle / plain torch
fp32 on the same seeded inputs.  Integer and index results are compared bit-exactly; fp32 results
within the tolerance written in each test (the north-star bound is 1e-4 relative on losses and
scores; most kernels are far inside it)."""
import math

import numpy as np
import pytest
import torch

from helpers import rel_err
import c2dsr_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from c2dsr_b200 import ops, _cabi
    return ops, _cabi


def _seq_with_pads(B, L, n_items, pad, gen, full_rows=0):
    """Left-padded sequences (>= 1 leading pad) with interior pads, plus `full_rows` rows without any pad."""
    seq = torch.randint(0, n_items, (B, L), generator=gen)
    lens = torch.randint(1, L, (B,), generator=gen)
    ar = torch.arange(L).unsqueeze(0)
    seq[ar < (L - lens).unsqueeze(1)] = pad
    seq[torch.rand(B, L, generator=gen) < 0.3] = pad
    seq[:, 0] = pad
    if full_rows:
        seq[:full_rows] = torch.randint(0, n_items, (full_rows, L), generator=gen)
    return seq


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,L,d,N", [(64, 15, 256, 3000), (1200, 15, 64, 500), (8, 10, 32, 121)])
def test_gather_fwd_bwd(B, L, d, N):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(B + d)
    pad = N - 1
    hi, E, P = torch.randn(N, d, generator=g), torch.randn(N, d, generator=g), torch.randn(L, d, generator=g)
    seq = _seq_with_pads(B, L, N - 1, pad, g)
    pos = torch.randint(0, L, (B, L), generator=g)
    dx = torch.randn(B, L, d, generator=g)
    scale = d ** 0.5
    # oracle: plain ops + autograd on CPU (padding_idx semantics: no direct gradient to the pad row)
    hi_c, E_c, P_c = (t.clone().requires_grad_(True) for t in (hi, E, P))
    direct = torch.where((seq == pad).unsqueeze(-1), E_c[seq].detach(), E_c[seq])
    x_ref = (hi_c[seq] + direct) * scale + P_c[pos]
    x_ref.backward(dx)

    hi_g, E_g, P_g = (t.to(DEV).requires_grad_(True) for t in (hi, E, P))
    x = ops.GatherFn.apply(hi_g, E_g, P_g, seq.to(DEV), pos.to(DEV), scale, pad, 0.0, 0, 0)
    assert torch.equal(x.cpu(), x_ref.detach())                      # same fp32 op order: bit-exact
    x.backward(dx.to(DEV))
    for got, ref, nm in ((hi_g.grad, hi_c.grad, "d_hi"), (E_g.grad, E_c.grad, "d_E"), (P_g.grad, P_c.grad, "d_P")):
        assert rel_err(got.cpu(), ref) < 1e-5, nm      # the pad row sums thousands of rows: order-dependent ulps
    assert float(E_g.grad[pad].abs().max()) == 0.0 and float(hi_g.grad[pad].abs().max()) > 0
    # determinism: a second backward gives identical bits (no float atomics)
    g1 = hi_g.grad.clone()
    hi_g.grad = None; E_g.grad = None; P_g.grad = None
    ops.GatherFn.apply(hi_g, E_g, P_g, seq.to(DEV), pos.to(DEV), scale, pad, 0.0, 0, 0).backward(dx.to(DEV))
    assert torch.equal(g1, hi_g.grad)


def test_gather_dropout_mask_consistency():
    """Forward keeps ~1-p of the entries scaled by 1/(1-p); backward regenerates the same mask."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(5)
    N, d, B, L, p = 200, 64, 256, 15, 0.25
    hi = torch.randn(N, d, generator=g).to(DEV); E = torch.randn(N, d, generator=g).to(DEV)
    P = torch.randn(L, d, generator=g).to(DEV).requires_grad_(True)
    seq = torch.randint(0, N, (B, L), generator=g).to(DEV)
    pos = torch.randint(0, L, (B, L), generator=g).to(DEV)
    x0 = ops.GatherFn.apply(hi, E, P, seq, pos, 1.0, N - 1, 0.0, 7, 3)
    x1 = ops.GatherFn.apply(hi, E, P, seq, pos, 1.0, N - 1, p, 7, 3)
    kept = x1 != 0
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.01
    assert torch.allclose(x1[kept], x0.detach()[kept] / (1 - p), rtol=1e-6)
    x1.backward(torch.ones_like(x1))
    # d_P[pos] = sum of mask/(1-p) over tokens with that position
    ref = torch.zeros_like(P).index_add_(0, pos.view(-1), kept.float().view(-1, d) / (1 - p))
    assert rel_err(P.grad.cpu(), ref.cpu()) < 1e-5


# ------------------------------------------------------------------------------------------------
def _random_csr(n, d, gen, heavy=700):
    deg = torch.randint(0, 8, (n,), generator=gen)
    deg[3] = heavy
    deg[5] = 0
    rows = torch.repeat_interleave(torch.arange(n), deg)
    cols = torch.randint(0, n, (rows.numel(),), generator=gen)
    vals = torch.rand(rows.numel(), generator=gen)
    return torch.sparse_coo_tensor(torch.stack((rows, cols)), vals, (n, n)).coalesce()


@pytest.mark.parametrize("d", [32, 256, 512])
def test_spmm_vs_sparse_mm(d):
    ops, _ = _ops()
    from c2dsr_b200.graph import CsrGraph
    g = torch.Generator().manual_seed(d)
    n = 2000
    A = _random_csr(n, d, g)
    X, Y, Z = (torch.randn(n, d, generator=g) for _ in range(3))
    G = CsrGraph(A, DEV)
    ref = 0.5 * torch.sparse.mm(A, X) + 0.25 * Y - 2.0 * Z
    out = ops.spmm(G.fwd, X.to(DEV), Y.to(DEV), Z.to(DEV), alpha=0.5, beta=0.25, gamma=-2.0)
    assert rel_err(out.cpu(), ref) < 2e-6
    ref_t = torch.sparse.mm(A.t().coalesce(), X)
    out_t = ops.spmm(G.bwd, X.to(DEV))
    assert rel_err(out_t.cpu(), ref_t) < 2e-6
    # in place on the addend
    Yd = Y.to(DEV).clone()
    ops.spmm(G.fwd, X.to(DEV), Y=Yd, out=Yd, alpha=1.0, beta=1.0)
    assert rel_err(Yd.cpu(), torch.sparse.mm(A, X) + Y) < 2e-6


@pytest.mark.parametrize("d,p", [(64, 0.0), (256, 0.25)])
def test_spmm_row_masks_are_exact(d, p):
    """out_row_needed: wanted rows bit-identical to the unmasked product, the others zero; x_row_nonzero:
    skipping rows of X that are all zero changes nothing (what a training step relies on: only the rows of the
    batch's items are propagated forward, only the rows it touched carry gradient backward)."""
    ops, _ = _ops()
    from c2dsr_b200.graph import CsrGraph
    g = torch.Generator().manual_seed(d)
    n = 3001
    A = _random_csr(n, d, g, heavy=300)
    G = CsrGraph(A, DEV)
    X, Y = torch.randn(n, d, generator=g).to(DEV), torch.randn(n, d, generator=g).to(DEV)
    ids = torch.randint(0, n, (400,), generator=g).to(DEV)
    need = ops.mark_rows(ids, n)
    assert int(need.sum()) == int(torch.unique(ids).numel())
    kw = dict(alpha=0.5, beta=0.5, drop_mode=1, p=p, seed=5, tag=9)
    full = ops.spmm(G.fwd, X, Y=Y, **kw)
    part = ops.spmm(G.fwd, X, Y=Y, out_need=need, **kw)
    sel = need.bool()
    assert torch.equal(part[sel], full[sel]) and float(part[~sel].abs().max()) == 0.0
    Xz = X.clone()
    Xz[~sel] = 0.0                                              # rows outside the mask are exactly zero
    kw2 = dict(alpha=0.5, beta=1.5, drop_mode=2, p=p, seed=5, tag=10)
    assert torch.equal(ops.spmm(G.bwd, Xz, Y=Xz, x_nz=need, **kw2), ops.spmm(G.bwd, Xz, Y=Xz, **kw2))


@pytest.mark.parametrize("n_gnn,p", [(1, 0.0), (2, 0.0), (3, 0.0), (1, 0.3), (2, 0.3)])
def test_gcn_forward_backward(n_gnn, p):
    """GCN mean-of-hops vs the oracle; with dropout the check is the adjoint identity
    <GCN(E), G> gradient == directional derivative (the map is linear in E for a fixed mask)."""
    ops, _ = _ops()
    from c2dsr_b200.graph import CsrGraph
    g = torch.Generator().manual_seed(11 + n_gnn)
    n, d = 1500, 64
    A = _random_csr(n, d, g, heavy=64)
    E = torch.randn(n, d, generator=g)
    Gm = torch.randn(n, d, generator=g)
    G = CsrGraph(A, DEV)
    E_g = E.to(DEV).requires_grad_(True)
    hi = ops.GCNFn.apply(E_g, G, n_gnn, p, 99, 1)
    hi.backward(Gm.to(DEV))
    if p == 0.0:
        E_c = E.clone().requires_grad_(True)
        ref = oracle.gcn(E_c, A, n_gnn)
        ref.backward(Gm)
        assert rel_err(hi.detach().cpu(), ref.detach()) < 2e-6
        assert rel_err(E_g.grad.cpu(), E_c.grad) < 5e-6
    else:
        V = torch.randn(n, d, generator=g).to(DEV)
        lin = ops.GCNFn.apply(V, G, n_gnn, p, 99, 1)                  # linear map applied to V (same mask)
        lhs = float((lin.double() * Gm.to(DEV).double()).sum())
        rhs = float((V.double() * E_g.grad.double()).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(130, 29, 77), (256, 384, 256), (64, 64, 4096), (1, 33, 500), (300, 1, 64)])
def test_gemm_all_layouts(ta, tb, M, N, K):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    opA, opB = (A.t() if ta else A).double(), (B.t() if tb else B).double()
    ref = torch.relu(0.5 * opA @ opB + bias.double()) + 2.0 * C0.double()
    Cd = C0.to(DEV).clone()
    ops.gemm(ta, tb, M, N, K, A.to(DEV), A.shape[1], B.to(DEV), B.shape[1], Cd, N, alpha=0.5, beta=2.0,
             bias=bias.to(DEV), act=1)
    assert rel_err(Cd.cpu(), ref) < 1e-5


# ------------------------------------------------------------------------------------------------
def _layer_norm_ref(s, w, b, eps=1e-8):
    mu = s.mean(-1, keepdim=True)
    var = ((s - mu) ** 2).mean(-1, keepdim=True)
    return (s - mu) / torch.sqrt(var + eps) * w + b


@pytest.mark.parametrize("d", [32, 256, 512])
def test_layernorm_family(d):
    _, cabi = _ops()
    g = torch.Generator().manual_seed(d)
    T = 700
    x, y, go = (torch.randn(T, d, generator=g) for _ in range(3))
    w, b = torch.randn(d, generator=g), torch.randn(d, generator=g)
    s_c = (x + y).requires_grad_(True)
    out_ref = _layer_norm_ref(s_c, w, b)
    w_c, b_c = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    _layer_norm_ref(s_c, w_c, b_c).backward(go)
    xd, yd, wd, bd, god = (t.to(DEV) for t in (x, y, w, b, go))
    s = torch.empty(T, d, device=DEV); out = torch.empty(T, d, device=DEV); st = torch.empty(T, 2, device=DEV)
    P = cabi.ptr
    cabi.call("c2dsr_add_ln_fwd", P(xd), P(yd), P(wd), P(bd), P(s), P(out), P(st), T, d, 1, 1e-8, 0.0, 0, 0,
              cabi.stream())
    assert rel_err(out.cpu(), out_ref.detach()) < 5e-6
    dx = torch.empty(T, d, device=DEV)
    cabi.call("c2dsr_ln_bwd", P(god), P(s), P(st), P(wd), P(dx), 0, T, d, cabi.stream())
    assert rel_err(dx.cpu(), s_c.grad) < 2e-5
    dw, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    cabi.call("c2dsr_ln_param_grad", P(god), P(s), P(st), P(dw), P(db), T, d, cabi.stream())
    assert rel_err(dw.cpu(), w_c.grad) < 2e-5 and rel_err(db.cpu(), b_c.grad) < 2e-5


# ------------------------------------------------------------------------------------------------
def _encoder_weights(d, n_layers, gen):
    W = {}
    for i in range(n_layers):
        p = f"enc.encoder.layers.{i}."
        W[p + "self_attn.in_proj_weight"] = torch.randn(3 * d, d, generator=gen) / math.sqrt(d)
        W[p + "self_attn.in_proj_bias"] = torch.randn(3 * d, generator=gen) * 0.1
        W[p + "self_attn.out_proj.weight"] = torch.randn(d, d, generator=gen) / math.sqrt(d)
        W[p + "self_attn.out_proj.bias"] = torch.randn(d, generator=gen) * 0.1
        for nm in ("linear1", "linear2"):
            W[p + nm + ".weight"] = torch.randn(d, d, generator=gen) / math.sqrt(d)
            W[p + nm + ".bias"] = torch.randn(d, generator=gen) * 0.1
        for nm in ("norm1", "norm2"):
            W[p + nm + ".weight"] = 1 + 0.1 * torch.randn(d, generator=gen)
            W[p + nm + ".bias"] = 0.1 * torch.randn(d, generator=gen)
    W["enc.encoder.norm.weight"] = 1 + 0.1 * torch.randn(d, generator=gen)
    W["enc.encoder.norm.bias"] = 0.1 * torch.randn(d, generator=gen)
    return W


def _weight_list(W, n_layers):
    out = []
    for i in range(n_layers):
        p = f"enc.encoder.layers.{i}."
        out += [W[p + k] for k in ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                                   "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight",
                                   "linear2.bias", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")]
    return out + [W["enc.encoder.norm.weight"], W["enc.encoder.norm.bias"]]


@pytest.mark.parametrize("B,L,d,H,n_layers,norm_first", [(48, 15, 256, 1, 1, False), (20, 10, 32, 2, 2, False),
                                                         (20, 10, 32, 4, 2, True), (6, 50, 128, 2, 1, False)])
@pytest.mark.parametrize("dense", [0, 3])
def test_encoder_fwd_bwd_vs_oracle(B, L, d, H, n_layers, norm_first, dense):
    """Includes rows with no pad at all (fully masked queries -> attention output 0)."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(B * L + d + H)
    pad = 999
    seq = _seq_with_pads(B, L, 900, pad, g, full_rows=2)
    x = torch.randn(B, L, d, generator=g)
    go = torch.randn(B, L, d, generator=g)
    W = _encoder_weights(d, n_layers, g)
    Wc = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    xc = x.clone().requires_grad_(True)
    ref = oracle.encoder(xc, seq, Wc, "enc", pad, H, n_layers, norm_first)
    ref.backward(go)
    wl = [t.to(DEV).requires_grad_(True) for t in _weight_list(W, n_layers)]
    xg = x.to(DEV).requires_grad_(True)
    out = ops.EncoderFn.apply(xg, seq.to(DEV), H, pad, norm_first, 0.0, 0, 0, dense, *wl)
    assert rel_err(out.detach().cpu(), ref.detach()) < 2e-5
    out.backward(go.to(DEV))
    # tensor-core dense layers (bf16 hi/lo split, ~1e-5 on pre-activations) flip a handful of ReLU units
    # per batch: the element-wise bound is looser there, the fp32 FFMA layers must match to 1e-4
    tol = 1e-4 if dense == 0 else 5e-3
    assert rel_err(xg.grad.cpu(), xc.grad) < tol
    for got, ref_t in zip(wl, _weight_list(Wc, n_layers)):
        scale = float(ref_t.grad.abs().max())
        assert float((got.grad.cpu() - ref_t.grad).abs().max()) <= tol * scale + 1e-6


@pytest.mark.parametrize("p", [0.0, 0.3])
def test_encoder_dropout_gradient_is_consistent(p):
    """With dropout on, fwd and bwd must use the same masks: compare the analytic directional
    derivative with a central finite difference of the (deterministic, seeded) forward.  p = 0
    calibrates the finite-difference error of the same check."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(3)
    B, L, d, H, nl, pad = 16, 10, 64, 2, 2, 999
    seq = _seq_with_pads(B, L, 900, pad, g).to(DEV)
    x = torch.randn(B, L, d, generator=g).to(DEV).requires_grad_(True)
    wl = [t.to(DEV) for t in _weight_list(_encoder_weights(d, nl, g), nl)]
    c = torch.randn(B, L, d, generator=g).to(DEV)
    v = torch.randn(B, L, d, generator=g).to(DEV)
    # fp32 FFMA projections for the finite difference: the bf16 hi/lo split of the tensor-core path has a
    # non-smooth ~1e-6 product error, i.e. a noise floor of O(1) in a difference quotient with eps ~ 1e-4
    f = lambda xx: float((ops.EncoderFn.apply(xx, seq, H, pad, False, p, 1234, 5, 0, *wl).double() * c.double()).sum())
    out = ops.EncoderFn.apply(x, seq, H, pad, False, p, 1234, 5, 0, *wl)
    (out * c).sum().backward()
    analytic = float((x.grad.double() * v.double()).sum())
    errs = []
    # the map has ReLU / attention kinks: the central difference only converges below eps ~ 3e-4
    # (checked with the fp64 oracle), where fp32 round-off of f adds ~0.1 absolute
    for eps in (3e-4, 1e-4):
        numeric = (f(x.detach() + eps * v) - f(x.detach() - eps * v)) / (2 * eps)
        errs.append(abs(analytic - numeric) - 0.03 * abs(numeric) - 0.15)
    assert min(errs) <= 0, (analytic, errs)
    # the tensor-core projections draw the same masks: same output and input gradient up to the split error
    x3 = x.detach().clone().requires_grad_(True)
    out3 = ops.EncoderFn.apply(x3, seq, H, pad, False, p, 1234, 5, 3, *wl)
    (out3 * c).sum().backward()
    assert rel_err(out3.detach().cpu(), out.detach().cpu()) < 1e-4
    assert rel_err(x3.grad.cpu(), x.grad.cpu()) < 1e-2
    if p > 0:                                                        # and dropout really is on
        out0 = ops.EncoderFn.apply(x.detach(), seq, H, pad, False, 0.0, 1234, 5, 3, *wl)
        assert rel_err(out.detach().cpu(), out0.cpu()) > 1e-2


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d_bias", [False, True])
def test_infomax_vs_oracle(d_bias):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(21)
    B, L, d = 96, 15, 128
    hs = [torch.randn(B, L, d, generator=g) for _ in range(5)]
    ma = (torch.rand(B, L, generator=g) < 0.4).long(); ma[:, -1] = 1
    mb = (torch.rand(B, L, generator=g) < 0.4).long(); mb[:, -2] = 1
    W = {"D_a.weight": torch.randn(1, d, d, generator=g) / d, "D_b.weight": torch.randn(1, d, d, generator=g) / d,
         "D_a.bias": torch.randn(1, generator=g), "D_b.bias": torch.randn(1, generator=g)}
    Wc = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    hc = [t.clone().requires_grad_(True) for t in hs]
    ref = oracle.infomax_loss(Wc, *hc, ma, mb, {"d_bias": d_bias})
    (ref * 0.3).backward()
    hg = [t.to(DEV).requires_grad_(True) for t in hs]
    Wg = {k: v.to(DEV).requires_grad_(True) for k, v in W.items()}
    loss = ops.InfomaxFn.apply(*hg, Wg["D_a.weight"], Wg["D_b.weight"], Wg["D_a.bias"] if d_bias else None,
                               Wg["D_b.bias"] if d_bias else None, ma.to(DEV), mb.to(DEV), 1.0 / B)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    (loss * 0.3).backward()
    for got, r in zip(hg, hc):
        assert rel_err(got.grad.cpu(), r.grad) < 5e-5
    keys = ["D_a.weight", "D_b.weight"] + (["D_a.bias", "D_b.bias"] if d_bias else [])
    for k in keys:
        assert rel_err(Wg[k].grad.cpu(), Wc[k].grad) < 5e-5, k


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,d", [(200, 777, 64), (640, 2903, 256)])
def test_score_ce_vs_oracle(M, N, d):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(M + N)
    H, Hp = torch.randn(M, d, generator=g), torch.randn(M, d, generator=g)
    Wt = torch.randn(N, d, generator=g) * 0.1
    b = torch.randn(N, generator=g) * 0.1
    wp, bp = torch.randn(1, d, generator=g) * 0.1, torch.randn(1, generator=g)
    gt = torch.randint(0, N + 1, (M,), generator=g)
    gt[:7] = N                                                        # ignored rows
    rs = torch.rand(M, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (H, Hp, Wt, b, wp, bp)]
    Hc, Hpc, Wc, bc, wpc, bpc = leaves
    z = torch.cat((Hc @ Wc.t() + bc, Hpc @ wpc.t() + bpc), -1)
    lse = torch.logsumexp(z, -1)
    picked = z.gather(1, gt.clamp(max=N).unsqueeze(1)).squeeze(1)
    ref = (((lse - picked) * (gt != N)) * rs).sum()
    (ref * 0.7).backward()
    dl = [t.to(DEV).requires_grad_(True) for t in (H, Hp, Wt, b, wp, bp)]
    loss = ops.ScoreCEFn.apply(*dl, gt.to(DEV), rs.to(DEV))
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))         # north-star bound is 1e-4
    (loss * 0.7).backward()
    for got, r, nm in zip(dl, leaves, ("dH", "dHpad", "dW", "db", "dwpad", "dbpad")):
        assert rel_err(got.grad.cpu(), r.grad) < 5e-5, nm


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_q,N,n_neg", [(64, 500, 99), (300, 29207, 999)])
def test_rank_counts_bit_exact(n_q, N, n_neg):
    """Integer contract: counts from given fp32 scores equal numpy exactly -- list mode, full mode,
    with ties, and summed over catalogue shards."""
    ops, _ = _ops()
    rng = np.random.default_rng(n_q)
    S = rng.standard_normal((n_q, N)).astype(np.float32)
    S[:, ::7] = np.round(S[:, ::7], 1)                                # plenty of exact ties
    gt = rng.integers(0, N, n_q)
    S[np.arange(n_q), gt] = np.round(S[np.arange(n_q), gt], 1)
    neg = np.stack([rng.choice(np.delete(np.arange(N), gt[i]), n_neg, replace=False) for i in range(n_q)])
    ref_list = oracle.rank_from_scores(S, gt, neg)
    ref_full = oracle.rank_from_scores(S, gt, None)
    lds = (N + 3) // 4 * 4
    Sd = torch.zeros(n_q, lds, device=DEV)
    Sd[:, :N] = torch.from_numpy(S).to(DEV)
    gtd, negd = torch.from_numpy(gt).to(DEV), torch.from_numpy(neg).to(DEV)
    s_gt = ops.pick_target(Sd, gtd, 0, N)
    assert np.array_equal(s_gt.cpu().numpy(), S[np.arange(n_q), gt])
    assert np.array_equal(ops.rank_from_scores(Sd, s_gt, gtd, negd, 0, N).cpu().numpy() + 1, ref_list)
    assert np.array_equal(ops.rank_from_scores(Sd, s_gt, gtd, None, 0, N).cpu().numpy() + 1, ref_full)
    # three uneven shards: partial counts add up, the target score comes from its owner
    bounds = [0, N // 3 + 1, N // 2, N]
    counts = torch.zeros(n_q, dtype=torch.int32, device=DEV)
    s_sum = torch.zeros(n_q, device=DEV)
    shards = []
    for n0, n1 in zip(bounds[:-1], bounds[1:]):
        sh = Sd[:, n0:n1].contiguous()
        shards.append((sh, n0, n1))
        s_sum += ops.pick_target(sh, gtd, n0, n1)
    assert torch.equal(s_sum, s_gt)
    for sh, n0, n1 in shards:
        ops.rank_from_scores(sh, s_sum, gtd, None, n0, n1, counts)
    assert np.array_equal(counts.cpu().numpy() + 1, ref_full)


def test_score_shard_matches_fp32_scores():
    ops, _ = _ops()
    g = torch.Generator().manual_seed(8)
    Q, W, b = torch.randn(70, 256, generator=g), torch.randn(1001, 256, generator=g) * 0.05, torch.randn(1001, generator=g)
    S = ops.score_shard(Q.to(DEV), W.to(DEV), b.to(DEV))[:, :1001]
    ref = Q.double() @ W.double().t() + b.double()
    assert float((S.cpu().double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


# ------------------------------------------------------------------------------------------------
def test_adamw_amsgrad_vs_oracle():
    from c2dsr_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(4)
    shapes = [(300, 64), (64,), (1, 64, 64), (5,)]
    ps = [torch.randn(*s, generator=g) for s in shapes]
    params = {str(i): p.clone() for i, p in enumerate(ps)}
    opt_ref = oracle.AdamWAmsgrad(params, lr=1e-2, weight_decay=5e-4)
    dev_p = [torch.nn.Parameter(p.to(DEV)) for p in ps]
    dead = torch.nn.Parameter(torch.ones(3, device=DEV))              # never gets a gradient
    opt = FusedAdamW(dev_p + [dead], lr=1e-2, weight_decay=5e-4, amsgrad=True)
    for step in range(4):
        grads = [torch.randn(*s, generator=g) * (10.0 ** -step) for s in shapes]
        opt_ref.add_grads({str(i): gr for i, gr in enumerate(grads)})   # accumulates like the reference (Q2)
        opt_ref.step()
        for p, gr in zip(dev_p, grads):
            p.grad = gr.to(DEV) if p.grad is None else p.grad + gr.to(DEV)
        opt.step()
    for i, p in enumerate(dev_p):
        assert rel_err(p.detach().cpu(), params[str(i)]) < 2e-6
    assert torch.equal(dead.detach().cpu(), torch.ones(3))


def test_adamw_device_step_state_and_gradient_sums():
    """The modes the Trainer uses: optimiser-owned per-epoch gradient sum (accumulate=True), step number and
    learning rate read from the device step state, and the flat sharded form (step_flat) -- all against the
    oracle's AdamW-amsgrad on accumulated gradients; a learning-rate change reaches the device."""
    from c2dsr_b200.optim import FusedAdamW
    from c2dsr_b200._cabi import call, ptr, query, stream
    g = torch.Generator().manual_seed(9)
    shapes = [(200, 32), (32,), (7,)]
    ps = [torch.randn(*s, generator=g) for s in shapes]
    params = {str(i): p.clone() for i, p in enumerate(ps)}
    opt_ref = oracle.AdamWAmsgrad(params, lr=1e-2, weight_decay=5e-4)
    dev_p = [torch.nn.Parameter(p.to(DEV)) for p in ps]
    opt = FusedAdamW(dev_p, lr=1e-2, weight_decay=5e-4, amsgrad=True, accumulate=True)
    state = torch.zeros(query("c2dsr_step_state_bytes"), dtype=torch.uint8, device=DEV)
    opt.attach_step_state(state)
    call("c2dsr_step_begin", ptr(state), 123, stream())
    # flat twin: the same parameters as one contiguous range
    flat = torch.cat([p.reshape(-1) for p in ps]).to(DEV)
    gflat = torch.empty_like(flat)
    opt2 = FusedAdamW([torch.nn.Parameter(flat.clone())], lr=1e-2, weight_decay=5e-4, amsgrad=True)
    state2 = torch.zeros_like(state)
    opt2.attach_step_state(state2)
    call("c2dsr_step_begin", ptr(state2), 123, stream())
    for step in range(5):
        if step == 3:                                              # scheduler-style change
            for o in (opt, opt2):
                o.param_groups[0]["lr"] = 5e-3
            opt_ref.lr = 5e-3
        grads = [torch.randn(*s, generator=g) * (10.0 ** -step) for s in shapes]
        opt_ref.add_grads({str(i): gr for i, gr in enumerate(grads)})
        opt_ref.step()
        for p, gr in zip(dev_p, grads):
            assert p.grad is None                                  # released by the previous step
            p.grad = gr.to(DEV)
        opt.step()
        call("c2dsr_step_begin", ptr(state), 123, stream())
        gflat.copy_(torch.cat([gr.reshape(-1) for gr in grads]).to(DEV))
        opt2.step_flat(flat, gflat)
        call("c2dsr_step_begin", ptr(state2), 123, stream())
    for i, p in enumerate(dev_p):
        assert rel_err(p.detach().cpu(), params[str(i)]) < 2e-6
        assert rel_err(opt.accumulated_grad(p).cpu(), opt_ref.acc[str(i)]) < 1e-6
    ref_flat = torch.cat([params[str(i)].reshape(-1) for i in range(len(ps))])
    assert rel_err(flat.cpu(), ref_flat) < 2e-6
    assert torch.equal(state.cpu()[:8], state2.cpu()[:8])          # both counted 6 steps


# ------------------------------------------------------------------------------------------------
# adjacency builder on the device vs the host path (utils/graph.py semantics)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,n_edges,seed", [(50, 0, 0), (50, 400, 1), (300, 5000, 2), (70000, 200000, 3)])
def test_graph_build_matches_host(n, n_edges, seed):
    from c2dsr_b200.graph import CsrGraph, _to_sparse, normalised_coo
    rng = np.random.default_rng(seed)
    # heavy-tailed sources / destinations with many duplicate pairs, some rows empty
    src = np.minimum((rng.pareto(1.2, n_edges) * 3).astype(np.int64), n - 1)
    dst = np.minimum((rng.pareto(1.0, n_edges) * 5).astype(np.int64), n - 1)
    edges = np.stack((src, dst), 1)
    got = CsrGraph.from_edges(edges, n, DEV)
    ref = CsrGraph(_to_sparse(normalised_coo(edges, n), n), DEV)
    for a, b in ((got.fwd, ref.fwd), (got.bwd, ref.bwd)):
        assert torch.equal(a[0].cpu(), b[0].cpu())                   # rowptr
        assert torch.equal(a[1].cpu(), b[1].cpu())                   # col (column-sorted rows)
        assert torch.equal(a[2].cpu(), b[2].cpu())                   # val, bit for bit
        assert (a[3] is None) == (b[3] is None) and (a[3] is None or torch.equal(a[3].cpu(), b[3].cpu()))
    assert got.nnz == ref.nnz
