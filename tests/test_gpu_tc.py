"""tcgen05 score path (K4b) on the B200: fused score GEMM + rank count vs the fp32 oracle.

Checks: (1) scores of the 3-pass bf16 split are fp32-grade (<= 1e-5 of the score scale; the
north-star bound is 1e-4), (2) the target score produced by the separate target pass is
bit-identical to the score the counting pass sees for that item, (3) counts are bit-exact given the
kernel's own scores (integer contract), (4) ranks agree with the fp32 FFMA path except where the
margin is below the arithmetic's resolution, (5) shard sums equal the unsharded counts."""
import numpy as np
import pytest
import torch

import helpers  # noqa: F401  (puts oracle/ on sys.path)
import c2dsr_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _problem(n_q, N, d, seed):
    g = torch.Generator().manual_seed(seed)
    Q = torch.randn(n_q, d, generator=g)
    W = torch.randn(N, d, generator=g) * 0.05
    b = torch.randn(N, generator=g) * 0.1
    gt = torch.randint(0, N, (n_q,), generator=g)
    return Q, W, b, gt


@pytest.mark.parametrize("n_q,N,d,passes", [(128, 128, 64, 3), (200, 1000, 256, 3), (2048, 29207, 256, 3),
                                            (300, 5000, 128, 1), (77, 333, 512, 3)])
def test_fused_score_count(n_q, N, d, passes):
    from c2dsr_b200 import ops
    Q, W, b, gt = _problem(n_q, N, d, n_q + N)
    Qd, Wd, bd, gtd = Q.to(DEV), W.to(DEV), b.to(DEV), gt.to(DEV)
    Wsplit = ops.split_bf16(Wd, passes == 3)
    counts, s_gt, S = ops.score_rank_tc(Qd, Wsplit, bd, gtd, 0, N, passes=passes, want_scores=True)
    S = S[:, :N]
    ref = Q.double() @ W.double().t() + b.double()
    scale = float(ref.abs().max())
    err = float((S.cpu().double() - ref).abs().max()) / scale
    assert err <= (1e-5 if passes == 3 else 1e-2), err
    # (2) self-consistent target score
    assert torch.equal(s_gt, S[torch.arange(n_q, device=DEV), gtd])
    # (3) integer contract on the kernel's own scores
    own = oracle.rank_from_scores(S.cpu().numpy(), gt.numpy(), None) - 1
    assert np.array_equal(counts.cpu().numpy(), own)
    # (4) against the fp32 FFMA path
    S32 = ops.score_shard(Qd, Wd, bd)
    c32 = ops.rank_from_scores(S32, ops.pick_target(S32, gtd, 0, N), gtd, None, 0, N)
    if passes == 3:
        # ranks may differ only by candidates whose fp64 margin to the target is below the resolution
        # of fp32 arithmetic (either path may flip those; the reference's own GEMV would too)
        diff = (counts - c32).abs().cpu()
        margin = (ref - ref[torch.arange(n_q), gt].unsqueeze(1)).abs()
        near = (margin < 4e-5 * scale).sum(1) - 1               # minus the target itself
        assert bool((diff <= near).all()), (int(diff.max()), float((diff > 0).float().mean()))


def test_sharded_counts_add_up():
    from c2dsr_b200 import ops
    n_q, N, d = 256, 10007, 256
    Q, W, b, gt = _problem(n_q, N, d, 5)
    Qd, Wd, bd, gtd = Q.to(DEV), W.to(DEV), b.to(DEV), gt.to(DEV)
    full, s_full, _ = ops.score_rank_tc(Qd, ops.split_bf16(Wd), bd, gtd, 0, N)
    bounds = [0, 1300, 5000, 5001, N]
    s_sum = torch.zeros(n_q, device=DEV)
    splits = []
    for n0, n1 in zip(bounds[:-1], bounds[1:]):
        ws = ops.split_bf16(Wd[n0:n1].contiguous())
        splits.append((ws, n0, n1))
        _, s, _ = ops.score_rank_tc(Qd, ws, bd[n0:n1].contiguous(), gtd, n0, n1, counts=torch.zeros(n_q, device=DEV, dtype=torch.int32))
        s_sum += s
    assert torch.equal(s_sum, s_full)                     # exactly one shard owns each target
    counts = torch.zeros(n_q, device=DEV, dtype=torch.int32)
    for ws, n0, n1 in splits:
        ops.score_rank_tc(Qd, ws, bd[n0:n1].contiguous(), gtd, n0, n1, s_gt=s_sum, counts=counts)
    assert torch.equal(counts, full)


@pytest.mark.parametrize("M,N,d,passes", [(200, 777, 64, 3), (640, 2903, 256, 3), (256, 50, 32, 3), (5120, 29207, 256, 3),
                                          (640, 2903, 256, 1)])
def test_score_ce_tc_vs_oracle(M, N, d, passes):
    """Cross-entropy over the full catalogue on tensor cores vs plain fp32 (trainer.py:131-152).
    Loss within 1e-5 (north star 1e-4); gradients within 1e-4 of their scale for the 3-pass split."""
    from c2dsr_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    H, Hp = torch.randn(M, d, generator=g), torch.randn(M, d, generator=g)
    Wt = torch.randn(N, d, generator=g) * 0.1
    b = torch.randn(N, generator=g) * 0.1
    wp, bp = torch.randn(1, d, generator=g) * 0.1, torch.randn(1, generator=g)
    gt = torch.randint(0, N + 1, (M,), generator=g)
    gt[:7] = N
    rs = torch.rand(M, generator=g) / M
    leaves = [t.clone().double().requires_grad_(True) for t in (H, Hp, Wt, b, wp, bp)]
    Hc, Hpc, Wc, bc, wpc, bpc = leaves
    z = torch.cat((Hc @ Wc.t() + bc, Hpc @ wpc.t() + bpc), -1)
    lse = torch.logsumexp(z, -1)
    picked = z.gather(1, gt.clamp(max=N).unsqueeze(1)).squeeze(1)
    ref = (((lse - picked) * (gt != N)) * rs.double()).sum()
    (ref * 0.7).backward()
    dl = [t.to(DEV).requires_grad_(True) for t in (H, Hp, Wt, b, wp, bp)]
    loss = ops.score_ce(*dl, gt.to(DEV), rs.to(DEV), path="tc", passes=passes)
    tol_l, tol_g = (1e-5, 1e-4) if passes == 3 else (2e-3, 3e-2)
    assert abs(float(loss.detach()) - float(ref)) <= tol_l * abs(float(ref))
    (loss * 0.7).backward()
    for got, r, nm in zip(dl, leaves, ("dH", "dHpad", "dW", "db", "dwpad", "dbpad")):
        err = float((got.grad.cpu().double() - r.grad).abs().max()) / float(r.grad.abs().max())
        assert err < tol_g, (nm, err)


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 136), (768, 256, 11520), (3840, 768, 256), (130, 70, 40)])
def test_gemm_tc_all_layouts(ta, tb, M, N, K):
    """tcgen05 GEMM with K-major and MN-major operands (transposed operands are read in place) vs fp64."""
    from c2dsr_b200 import ops
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K + ta * 2 + tb)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    opA, opB = (A.t() if ta else A).double(), (B.t() if tb else B).double()
    ref = torch.relu(opA @ opB + bias.double()) + C0.double()
    Cd = C0.to(DEV).clone()
    ops.gemm_tc(ta, tb, M, N, K, A.to(DEV), A.shape[1], B.to(DEV), B.shape[1], Cd, N, beta=1.0, bias=bias.to(DEV), act=1)
    err = float((Cd.cpu().double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 2e-5, err


@pytest.mark.parametrize("frac", [1.0, 0.55, 0.0])
def test_domain_loss_rows_vs_torch(frac):
    """ops.DomainLossTcFn (row partition + emit, tcgen05 logits / CE, backward through the inverse permutation)
    against plain torch on the same tensors (trainer.py:122-152 written out): loss and all gradients.  ``frac`` of
    the valid rows go through the GEMMs: 1.0 = exactly the rows that carry a target, 0.55 = fewer (the dropped rows
    then contribute nothing, as documented for M), 0.0 = none."""
    from c2dsr_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, L, R, d, N = 24, 15, 10, 64, 777
    hs = torch.randn(B, L, d, generator=g).to(DEV).requires_grad_(True)
    hd = torch.randn(B, L, d, generator=g).to(DEV).requires_grad_(True)
    W = (0.1 * torch.randn(N, d, generator=g)).to(DEV).requires_grad_(True)
    b = (0.1 * torch.randn(N, generator=g)).to(DEV).requires_grad_(True)
    wpad = (0.1 * torch.randn(1, d, generator=g)).to(DEV).requires_grad_(True)
    bpad = torch.randn(1, generator=g).to(DEV).requires_grad_(True)

    def targets():
        t = torch.randint(0, N, (B, L), generator=g)
        t[torch.rand(B, L, generator=g) < 0.6] = N                     # ignore class
        return t.to(DEV)
    gs, gd = targets(), targets()
    valid = int((gs[:, -R:] != N).sum() + (gd[:, -R:] != N).sum())
    n_dom = (gd[:, -R:] != N).sum().float()
    w_share = torch.tensor(1.0 / (R * B), device=DEV)
    M = int(round(frac * valid))
    loss = ops.DomainLossTcFn.apply(hs, hd, W, b, wpad, bpad, gs, gd, w_share, n_dom, R, M, 3)
    loss.backward()
    got = [loss.detach()] + [t.grad.clone() for t in (hs, hd, W, b, wpad, bpad)]
    for t in (hs, hd, W, b, wpad, bpad):
        t.grad = None
    # torch restatement: virtual rows in the same order, first M valid ones kept
    hsr, hdr = hs[:, -R:].reshape(-1, d), hd[:, -R:].reshape(-1, d)
    H = torch.cat((hsr, hsr + hdr))
    Hp = torch.cat((hsr, hdr))
    gt = torch.cat((gs[:, -R:].reshape(-1), gd[:, -R:].reshape(-1)))
    w = torch.cat((w_share.expand(B * R), (1.0 / n_dom).expand(B * R)))
    keep = torch.nonzero(gt != N).view(-1)[:M]
    Z = torch.cat((H[keep] @ W.t() + b, Hp[keep] @ wpad.t() + bpad), 1)
    ref = (torch.nn.functional.cross_entropy(Z, gt[keep], reduction="none") * w[keep]).sum() if M else Z.sum() * 0
    ref.backward()
    want = [ref.detach()] + [t.grad if t.grad is not None else torch.zeros_like(t) for t in (hs, hd, W, b, wpad, bpad)]
    for a, c, nm in zip(got, want, ("loss", "d_h_share", "d_h_dom", "dW", "db", "dwpad", "dbpad")):
        scale = float(c.abs().max()) + 1e-12
        assert float((a - c).abs().max()) <= 2e-4 * scale + 1e-7, (nm, float((a - c).abs().max()), scale)
