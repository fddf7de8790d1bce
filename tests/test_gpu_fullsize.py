"""End-to-end parity at BASELINE.json sizes (configs[1] Food-Kitchen and configs[2] Movie-Book shapes, d = 256,
L = 15, batch 256 / 2 048) against the CPU oracle on the same seeded synthetic inputs and the same initial weights,
through the entry points bench.py times: ``Trainer.train_step`` (eager, eager, capture, replay) and
``Trainer.evaluate_batch`` (full catalogue).  Dropouts 0 (RNG streams cannot match).  Bars: losses 1e-4 relative,
full-catalogue ranks equal up to near-ties the fp32 reference itself decides by < 1e-6 margins,
Recall / MRR / NDCG @ {5, 20} within 1e-3 absolute (about 1 000 queries per domain)."""
import os
import sys

import numpy as np
import pytest
import torch

import helpers  # noqa: F401
import c2dsr_oracle as oracle

sys.path.insert(0, helpers.ROOT)
pytestmark = pytest.mark.gpu
DEV = "cuda"


class Quiet:
    def log_train(self, *a):
        pass

    def log_msg(self, *a):
        pass


def _build(shape, n_train_batches, n_eval_batches):
    import bench
    from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
    from c2dsr_b200.trainer import Trainer
    hp = bench.hyper(bench.WORKLOADS[shape], 0.0, torch.device(DEV))
    adj, fields, ev = bench.make_workload(hp, n_train_batches, n_eval_batches, seed=0)
    B, Bq, L = hp.batch_size, hp.batch_size_eval, hp.len_max
    torch.manual_seed(hp.seed)
    ds = CDSRDataset.from_fields([fields[:, i] for i in range(14)], "train", L)
    loader = BatchLoader(ds, B, len_rec=hp.len_rec, ignore=(hp.n_item_a, hp.n_item_b))
    tr = Trainer.from_parts(hp, Quiet(), (loader, None, None), adj[0], adj[1])
    h = {k: v for k, v in vars(hp).items() if isinstance(v, (int, float, bool, str))}
    state = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    torch.set_num_threads(os.cpu_count() or 1)
    otr = oracle.OracleTrainer(state, adj[0].coalesce(), adj[1].coalesce(), h)
    return hp, tr, otr, bench.train_batches(fields, B, False), bench.eval_batches(ev, Bq, False)


def _check_eval(tr, otr, ebatch):
    from c2dsr_b200.metrics import cal_metrics
    tr.model.eval()
    with torch.no_grad():
        tr.model.convolve_graph()
        ra, rb = tr.evaluate_batch(ebatch)
    otr.convolve_graph()
    oa, ob = otr.evaluate_batch(ebatch, full_catalog=True)
    assert len(ra) == len(oa) and len(rb) == len(ob) and len(ra) + len(rb) == ebatch[0].shape[0]
    diff = np.abs(np.asarray(ra + rb) - np.asarray(oa + ob))
    # ranks among 29k-64k candidates: a near-tie (score margin below fp32 resolution of the dot product) moves a
    # rank by one place; nothing may move by more than a few places and the bulk must be identical
    assert diff.max() <= 3 and (diff > 0).mean() < 0.10, (int(diff.max()), float((diff > 0).mean()))
    for got, ref in ((ra, oa), (rb, ob)):
        assert np.abs(np.asarray(cal_metrics(got)) - np.asarray(cal_metrics(ref))).max() <= 1e-3
    return float((diff > 0).mean())


def test_food_kitchen_shape_train_step_and_eval_match_oracle():
    """BASELINE configs[1]: 4 x Trainer.train_step (two eager, the capture, one replay) and one evaluation batch of
    2 048 queries over the full catalogue vs the oracle; the evaluation runs first, on the shared initial weights."""
    hp, tr, otr, tb, eb = _build("fk", 4, 1)
    init = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    _check_eval(tr, otr, eb[0])
    tr.model.train()
    tr.optimizer.zero_grad()
    otr.zero_grad()
    for s in range(4):
        got = [float(x) for x in tr.train_step(tb[s])]
        ref = [float(x) for x in otr.train_batch(tb[s], training=True)]
        np.testing.assert_allclose(got, ref, rtol=1e-4, err_msg=f"step {s}")
    assert tr._graphs, "the third step should have been captured"
    # weights after 4 accumulated-gradient AdamW steps.  AdamW moves every element by ~lr * g / sqrt(v) per step,
    # i.e. by +-lr whatever the size of g, so an element whose gradient is rounding noise (|g| ~ 1e-9) may step
    # the other way: compare the mean deviation with the mean distance travelled instead of the worst element
    for k in ("classifier_a.weight", "classifier_b.weight", "classifier_b.bias", "embed_i.weight", "embed_i_a.weight"):
        got, ref, start = tr.model.state_dict()[k].cpu(), otr.W[k].detach(), init[k]
        moved = float((ref - start).abs().mean())
        dev = float((got - ref).abs().mean())
        assert dev < 2e-2 * moved, (k, dev, moved, helpers.rel_err(got, ref))


def test_movie_book_shape_eval_matches_oracle():
    """BASELINE configs[2] catalogue (36 845 + 63 937 items): one full-catalogue evaluation batch of 2 048 queries
    and one training step vs the oracle."""
    hp, tr, otr, tb, eb = _build("mb", 1, 1)
    _check_eval(tr, otr, eb[0])
    tr.model.train()
    tr.optimizer.zero_grad()
    otr.zero_grad()
    got = [float(x) for x in tr.train_step(tb[0])]
    ref = [float(x) for x in otr.train_batch(tb[0], training=True)]
    np.testing.assert_allclose(got, ref, rtol=1e-4)
