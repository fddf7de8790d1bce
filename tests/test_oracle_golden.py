"""Pin the CPU oracle (oracle/c2dsr_oracle.py) to outputs of the reference itself.

The golden ``.npz`` files were produced by tests/golden/make_golden.py running the
unmodified reference on CPU.  Tolerances: losses/activations 1e-5 relative (fp32, different
op order), ranks and metrics exact.
"""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_NAMES, Golden, rel_err
import c2dsr_oracle as oracle


@pytest.fixture(scope="module", params=GOLDEN_NAMES)
def g(request):
    return Golden(request.param)


def test_graph_matches_reference(g):
    (srow, scol, sval), (prow, pcol, pval) = oracle.build_adjacency(g.raw("train"), g.hp["n_item_a"], g.hp["n_item"])
    for nm, (r, c, v) in (("share", (srow, scol, sval)), ("spec", (prow, pcol, pval))):
        assert np.array_equal(r, g.z[f"adj_{nm}_row"]) and np.array_equal(c, g.z[f"adj_{nm}_col"])
        np.testing.assert_allclose(v, g.z[f"adj_{nm}_val"], rtol=1e-7)


def test_forward_activations(g):
    W = g.group("init")
    if g.hp.get("shared_item_embed"):
        assert torch.equal(W["embed_i.weight"], W["embed_i_a.weight"])
    hi = oracle.convolve_graph(W, g.adj("share"), g.adj("spec"), g.hp, training=False)
    for got, key in zip(hi, ("hi_share", "hi_a", "hi_b")):
        assert rel_err(got, g.z["step0/" + key]) < 1e-6
    b = g.train_batch(0)
    hs, hx, hy = oracle.forward(W, hi, *b[:6], g.hp, training=False)
    for got, key in ((hs, "h_share"), (hx, "hx"), (hy, "hy")):
        assert rel_err(got, g.z["step0/" + key]) < 2e-5, key
    assert rel_err(oracle.forward_share(W, hi, b[12], b[3], g.hp, False), g.z["step0/h_neg_a"]) < 2e-5


def test_training_steps_losses_grads_weights(g):
    tr = oracle.OracleTrainer(g.group("init"), g.adj("share"), g.adj("spec"), g.hp)
    tr.zero_grad()
    ref_losses = g.z["losses"]
    for s in range(len(ref_losses)):
        out = tr.train_batch(g.train_batch(s), training=True)       # dropouts are 0 in the fixtures
        np.testing.assert_allclose([float(x) for x in out], ref_losses[s], rtol=2e-6)
        if s == 0:
            grads = g.group("grad0")
            assert len(grads) > 10
            for k, ref in grads.items():
                assert k in tr.opt.acc, k
                assert rel_err(tr.opt.acc[k], ref) < 5e-5, k
            # dead prototype layers never receive a gradient
            assert not any(".encoder_layer." in k for k in tr.opt.acc)
    final = g.group("final")
    d, lr, n = g.hp["d_latent"], g.hp["lr"], len(ref_losses)
    for k, ref in final.items():
        if k.endswith("attn_mask"):
            continue
        got = tr.W[k].detach()
        if "in_proj" in k:
            # With dropout 0 every attended key of a sequence is the same pad token, so the q/k
            # gradients are mathematically 0 and numerically ~1e-9 rounding noise; AdamW (eps 1e-8)
            # turns that noise into O(lr) steps that no implementation can reproduce.  Pin the
            # value rows exactly and bound the q/k rows by the largest possible AdamW walk.
            assert rel_err(got[2 * d:], ref[2 * d:]) < 2e-5, k
            assert float((got[:2 * d] - ref[:2 * d]).abs().max()) <= 2 * n * lr, k
            continue
        assert rel_err(got, ref) < 2e-5, k


def test_eval_ranks_and_metrics_exact(g):
    tr = oracle.OracleTrainer(g.group("final"), g.adj("share"), g.adj("spec"), g.hp)
    tr.convolve_graph()
    ra, rb = tr.evaluate_batch(g.eval_batch("val"))
    assert ra == g.z["eval/rank_a"].tolist() and rb == g.z["eval/rank_b"].tolist()
    np.testing.assert_allclose(oracle.cal_score(ra, rb, [0.1124, 0.0865, 0.0574, 0.0416]), g.z["eval/score"],
                               rtol=1e-12)
    # batched integer-exact counting form agrees with the per-sample loop
    q = tr.eval_queries(g.eval_batch("val"))
    b = g.eval_batch("val")
    dom, gt, neg = b[8].view(-1), b[9].view(-1), b[10]
    for d, (Wk, bk, ref) in enumerate((("classifier_a.weight", "classifier_a.bias", ra),
                                       ("classifier_b.weight", "classifier_b.bias", rb))):
        sel = dom == d
        s = (q[sel] @ tr.W[Wk].detach().t() + tr.W[bk].detach()).numpy()
        got = oracle.rank_from_scores(s, gt[sel].numpy(), neg[sel].numpy())
        assert sum(abs(int(a) - int(c)) for a, c in zip(got, ref)) <= 1   # GEMM vs GEMV rounding may flip one tie
        full = oracle.rank_from_scores(s, gt[sel].numpy(), None)
        assert (full >= 1).all() and (full <= s.shape[1]).all()


def test_metrics_known_values():
    m = oracle.cal_metrics([1, 3, 6, 21])
    assert m[0] == 0.5 and m[1] == 0.75
    assert abs(m[2] - (1 + 1 / 3) / 4) < 1e-15 and abs(m[4] - (1 + 0.5) / 4) < 1e-15
    with pytest.raises(ZeroDivisionError):
        oracle.cal_metrics([])
