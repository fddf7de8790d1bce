"""End-to-end parity on the B200 through the reference-shaped API (Trainer / C2DSR), against the
golden vectors produced by the reference itself (tests/golden/make_golden.py) and against the CPU
oracle at sizes it finishes in seconds.  Tolerances (north star): losses and scores 1e-4 relative,
ranks exact given identical scores, Recall/MRR/NDCG within 1e-3 absolute."""
import argparse

import numpy as np
import pytest
import torch

from helpers import GOLDEN_NAMES, Golden, rel_err
import c2dsr_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


class QuietNoter:
    def log_train(self, *a):
        pass

    def log_msg(self, *a):
        pass


def _trainer_from_golden(g, state="init", **over):
    from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
    from c2dsr_b200.trainer import Trainer
    hp = dict(g.hp)
    hp.update(over)
    args = argparse.Namespace(**hp)
    args.device = torch.device(DEV)
    z = g.z
    train = CDSRDataset.from_fields([z["train_fields"][:, i] for i in range(14)], "train", hp["len_max"])
    evals = {}
    for mode in ("val", "test"):
        six, four, neg = z[f"{mode}_six"], z[f"{mode}_four"], z[f"{mode}_neg"]
        evals[mode] = CDSRDataset.from_fields([six[:, i] for i in range(6)] + [four[:, i:i + 1] for i in range(4)]
                                              + [neg], mode, hp["len_max"])
    loaders = (BatchLoader(train, hp["batch_size"]), BatchLoader(evals["val"], hp["batch_size_eval"]),
               BatchLoader(evals["test"], hp["batch_size_eval"]))
    torch.manual_seed(hp["seed"])
    tr = Trainer.from_parts(args, QuietNoter(), loaders, g.adj("share"), g.adj("spec"))
    tr.model.load_state_dict({k: v.to(DEV) for k, v in g.group(state).items()})
    return tr


@pytest.mark.parametrize("dense", [0, 3])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_training_steps_match_reference(name, dense):
    """dense = 0: encoder projections in fp32 FFMA; 3: on tcgen05 with the bf16 hi/lo split (the default)."""
    g = Golden(name)
    tr = _trainer_from_golden(g, encoder_tc_passes=dense)
    assert tr.model.attn_share.dense_passes == dense
    tr.model.train()                                   # dropouts are 0 in the fixtures
    tr.optimizer.zero_grad()
    ref_losses = g.z["losses"]
    for s in range(len(ref_losses)):
        tr.model.convolve_graph()
        if s == 0:
            with torch.no_grad():
                b = tuple(x.to(DEV) for x in g.train_batch(0))
                hs, hx, hy = tr.model(*b[:6])
                for got, key in ((hs, "h_share"), (hx, "hx"), (hy, "hy")):
                    assert rel_err(got.cpu(), g.z["step0/" + key]) < 1e-4, key
                assert rel_err(tr.model.hi_share.cpu(), g.z["step0/hi_share"]) < 1e-5
                assert rel_err(tr.model.forward_share(b[12], b[3]).cpu(), g.z["step0/h_neg_a"]) < 1e-4
        out = tr.train_batch(g.train_batch(s))
        np.testing.assert_allclose([float(x) for x in out], ref_losses[s], rtol=1e-4)
        if s == 0:
            grads = g.group("grad0")
            named = dict(tr.model.named_parameters())
            acc = tr.optimizer.accumulated_grad       # what p.grad holds in the reference (Q2: summed per epoch)
            for k, ref in grads.items():
                assert acc(named[k]) is not None, k
                if "in_proj" in k:                     # q/k rows: rounding noise only (see test_oracle_golden)
                    d = g.hp["d_latent"]
                    assert rel_err(acc(named[k])[2 * d:].cpu(), ref[2 * d:]) < 1e-3, k
                    continue
                assert rel_err(acc(named[k]).cpu(), ref) < 1e-3, k
            assert all(acc(p) is None for k, p in named.items() if ".encoder_layer." in k)   # Q3
    final = g.group("final")
    d, lr, n = g.hp["d_latent"], g.hp["lr"], len(ref_losses)
    # AdamW's update is g / sqrt(v), so a parameter follows the *relative* error of its gradient element.  A
    # ReLU unit flipped by the ~1e-6 product error of the tensor-core split changes the few-term sums behind
    # a bias gradient of these tiny fixtures (d = 32, 320 tokens) by percents; losses stay within 1e-4 (above)
    # and the fp32 FFMA mode keeps the 1e-3 bar on every parameter
    w_tol = 1e-3 if dense == 0 else 3e-2
    for k, p in tr.model.state_dict().items():
        if k.endswith("attn_mask"):
            continue
        ref, got = final[k], p.cpu()
        if "in_proj" in k:
            assert rel_err(got[2 * d:], ref[2 * d:]) < w_tol, k
            assert float((got[:2 * d] - ref[:2 * d]).abs().max()) <= 2 * n * lr, k
            continue
        assert rel_err(got, ref) < w_tol, k


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_eval_ranks_and_metrics_match_reference(name):
    from c2dsr_b200.metrics import cal_score
    g = Golden(name)
    tr = _trainer_from_golden(g, state="final")
    tr.model.eval()
    with torch.no_grad():
        tr.model.convolve_graph()
        batch = g.eval_batch("val")
        ra, rb = tr.evaluate_batch(batch)
        hs, hx, hy = tr.model(*(x.to(DEV) for x in batch[:6]))
    if "eval/h_share" in g.z.files:
        for got, key in ((hs, "h_share"), (hx, "hx"), (hy, "hy")):
            assert rel_err(got.cpu(), g.z["eval/" + key]) < 1e-4, key
    ref_a, ref_b = g.z["eval/rank_a"].tolist(), g.z["eval/rank_b"].tolist()
    assert len(ra) == len(ref_a) and len(rb) == len(ref_b)
    # identical up to comparisons the reference itself decides by < 1e-6 score margins (a handful among the
    # 2 400 queries of mid_default, at most one in the 48-query fixtures)
    n_q = len(ra) + len(rb)
    assert sum(abs(x - y) for x, y in zip(ra + rb, ref_a + ref_b)) <= max(1, n_q // 400)
    got = cal_score(ra, rb, [0.1124, 0.0865, 0.0574, 0.0416])
    # the north-star bar: Recall / MRR / NDCG @ {5, 20} within 1e-3 absolute.  mid_default has > 1 000 queries per
    # domain, so the bar is a real one there; in the 48-query fixtures one flipped near-tie moves a metric by
    # 1 / n, so there the ranks themselves must be identical for the bar to be applied
    if min(len(ra), len(rb)) >= 1000 or ra + rb == ref_a + ref_b:
        assert np.abs(np.asarray(got[1:]) - g.z["eval/score"][1:]).max() <= 1e-3
    # full-catalogue mode against the oracle on the same weights
    tr.full_catalog = True
    fa, fb = tr.evaluate_batch(batch)
    otr = oracle.OracleTrainer(g.group("final"), g.adj("share"), g.adj("spec"), g.hp)
    otr.convolve_graph()
    oa, ob = otr.evaluate_batch(batch, full_catalog=True)
    assert sum(abs(x - y) for x, y in zip(fa + fb, oa + ob)) <= 1
    assert all(f >= r for f, r in zip(fa + fb, ra + rb))           # more candidates can only push the rank up


@pytest.mark.parametrize("name", ["mid_default", "tiny_deep", "tiny_prenorm_shared"])
def test_train_step_matches_reference(name):
    """The TIMED entry point (bench.py times Trainer.train_step) against the reference's golden losses, first-step
    gradients and final weights: lazy row-masked propagation inside the branch set, capacity-padded loss rows,
    two eager steps, the capture, then graph replays (mid_default has 8 steps)."""
    g = Golden(name)
    tr = _trainer_from_golden(g, encoder_tc_passes=0)
    tr.model.train()
    tr.optimizer.zero_grad()
    ref_losses = g.z["losses"]
    for s in range(len(ref_losses)):
        out = tr.train_step(g.train_batch(s))
        np.testing.assert_allclose([float(x) for x in out], ref_losses[s], rtol=1e-4, err_msg=f"step {s}")
        if s == 0:
            named = dict(tr.model.named_parameters())
            for k, ref in g.group("grad0").items():
                got = tr.optimizer.accumulated_grad(named[k])
                assert got is not None, k
                d = g.hp["d_latent"]
                sl = slice(2 * d, None) if "in_proj" in k else slice(None)
                assert rel_err(got[sl].cpu(), ref[sl]) < 1e-3, k
    assert bool(tr._graphs) == (len(ref_losses) >= 3)                  # the third step is the capture
    if len(ref_losses) > 3:
        assert next(iter(tr._graphs.values()))["overflows"] == 0        # ... and the later ones were replays
    # the propagated tables the step leaves behind are the reference's (Q13), on the rows the batch touched
    final = g.group("final")
    d, lr, n = g.hp["d_latent"], g.hp["lr"], len(ref_losses)
    for k, p in tr.model.state_dict().items():
        if k.endswith("attn_mask"):
            continue
        ref, got = final[k], p.cpu()
        if "in_proj" in k:
            assert rel_err(got[2 * d:], ref[2 * d:]) < 1e-3, k
            assert float((got[:2 * d] - ref[:2 * d]).abs().max()) <= 2 * n * lr, k
            continue
        assert rel_err(got, ref) < 1e-3, k


@pytest.mark.parametrize("name", ["mid_default", "tiny_prenorm_shared"])
def test_early_optimiser_steps_are_bit_identical(name):
    """Updating the large tensors beside the backward (FusedAdamW.enable_early: classifier matrices from their
    post-accumulate-grad hook, embedding tables from inside their branch's backward, gradients in persistent sinks)
    is the same arithmetic in a different launch order: after eager steps, the capture and replays -- with dropout on
    -- every parameter and every per-epoch gradient sum equals the end-of-backward update bit for bit."""
    g = Golden(name)
    states = []
    for early in (1, 0):
        tr = _trainer_from_golden(g, early_adam=early, dropout_gnn=0.2, dropout_attn=0.2)
        assert bool(getattr(tr.optimizer, "_early", None)) == bool(early)
        tr.model.train()
        tr.optimizer.zero_grad()
        losses = [[float(x) for x in tr.train_step(g.train_batch(s))] for s in range(len(g.z["losses"]))]
        named = dict(tr.model.named_parameters())
        states.append((losses, {k: p.detach().cpu().clone() for k, p in named.items()},
                       {k: tr.optimizer.accumulated_grad(p).cpu().clone() for k, p in named.items()
                        if tr.optimizer.accumulated_grad(p) is not None}))
    (l1, p1, a1), (l0, p0, a0) = states
    assert l1 == l0
    assert p1.keys() == p0.keys() and a1.keys() == a0.keys()
    for k in p1:
        assert torch.equal(p1[k], p0[k]), k
    for k in a1:
        assert torch.equal(a1[k], a0[k]), k


def test_run_epoch_and_run_test_contract():
    """Trainer.run_epoch / run_test return two python lists of int ranks (trainer.py:40-83)."""
    g = Golden("tiny_default")
    tr = _trainer_from_golden(g, dropout_gnn=0.2, dropout_attn=0.2)       # reference default dropouts
    va, vb = tr.run_epoch()
    ta, tb = tr.run_test()
    assert len(va) + len(vb) == len(g.z["val_four"]) and len(ta) + len(tb) == len(g.z["test_four"])
    assert all(isinstance(r, int) and 1 <= r <= g.hp["n_neg_sample"] + 1 for r in va + vb + ta + tb)
    # training moved the weights and the loss is finite
    assert not torch.equal(tr.model.classifier_a.weight.cpu(), g.group("init")["classifier_a.weight"])
    assert all(torch.isfinite(p).all() for p in tr.model.parameters())


@pytest.mark.parametrize("p", [0.0, 0.2])
def test_captured_step_equals_eager_step(p):
    """Trainer.train_step replays a CUDA graph from the third step on.  Eager and replayed steps read the
    same device step state (optimiser step number, dropout key words), so with the same batches the two
    trainers must produce the same losses and weights (up to the summation order of the loss GEMMs, whose
    row capacity differs) -- with dropout on as well as off."""
    from c2dsr_b200 import _cabi
    g = Golden("tiny_default")
    runs = []
    for graph in (False, True):
        tr = _trainer_from_golden(g, dropout_gnn=p, dropout_attn=p, cuda_graph=graph)
        tr.model.train()
        tr.optimizer.zero_grad()
        l0 = _cabi.launch_count()
        losses = []
        for s in range(7):
            b = tuple(x.to(DEV) for x in g.train_batch(s % 3))
            losses.append([float(x) for x in tr.train_step(b)])
        assert bool(tr._graphs) == graph
        runs.append((losses, {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()},
                     _cabi.launch_count() - l0))
    (le, we, ne), (lg, wg, ng) = runs
    np.testing.assert_allclose(lg, le, rtol=2e-5)
    for k in we:
        if not k.endswith("attn_mask"):
            assert rel_err(wg[k], we[k]) < 1e-4, k
    assert ng > 0.9 * ne                       # replayed launches are counted


def test_training_with_dropout_learns():
    """Dropout path sanity at the reference's default rates: loss decreases over a few epochs."""
    g = Golden("tiny_default")
    tr = _trainer_from_golden(g, dropout_gnn=0.2, dropout_attn=0.2, lr=5e-3)
    tr.model.train()
    first = last = None
    for epoch in range(6):
        tr.optimizer.zero_grad()
        tot = 0.0
        for batch in tr.trainloader:
            tr.model.convolve_graph()
            tot += float(tr.train_batch(batch)[0])
        first = tot if first is None else first
        last = tot
    assert last < first


def test_full_size_food_kitchen_eval_properties():
    """BASELINE config 2 size (29 207 + 34 886 items, d = 256, 2 048 queries): size-independent
    properties of the full-catalogue ranking -- shard sums equal the unsharded counts, list-mode rank
    <= full rank, permuting the catalogue leaves every rank unchanged."""
    from c2dsr_b200 import ops
    gen = torch.Generator().manual_seed(0)
    n_q, N, d = 2048, 34886, 256
    Q = torch.randn(n_q, d, generator=gen).to(DEV)
    W = (torch.randn(N, d, generator=gen) * 0.01).to(DEV)
    b = torch.zeros(N, device=DEV)
    gt = torch.randint(0, N, (n_q,), generator=gen).to(DEV)
    S = ops.score_shard(Q, W, b)
    s_gt = ops.pick_target(S, gt, 0, N)
    full = ops.rank_from_scores(S, s_gt, gt, None, 0, N)
    ref = (S[:, :N] > s_gt[:, None]).sum(1).to(torch.int32)
    assert torch.equal(full, ref)
    neg = torch.randint(0, N - 1, (n_q, 999), generator=gen).to(DEV)
    neg = neg + (neg >= gt[:, None]).long()
    assert bool((ops.rank_from_scores(S, s_gt, gt, neg, 0, N) <= full).all())
    perm = torch.randperm(N, generator=gen).to(DEV)
    inv = torch.empty_like(perm); inv[perm] = torch.arange(N, device=DEV)
    S2 = ops.score_shard(Q, W[perm].contiguous(), b)
    gt2 = inv[gt]
    assert torch.equal(ops.rank_from_scores(S2, ops.pick_target(S2, gt2, 0, N), gt2, None, 0, N), full)
    counts = torch.zeros(n_q, dtype=torch.int32, device=DEV)
    for r in range(8):                                                # 8 catalogue shards, as on 8 GPUs
        n0, n1 = (N + 7) // 8 * r, min((N + 7) // 8 * (r + 1), N)
        Sr = ops.score_shard(Q, W[n0:n1], b[n0:n1])
        ops.rank_from_scores(Sr, s_gt, gt, None, n0, n1, counts)
    assert torch.equal(counts, full)


def test_bench_json_contract_on_gpu():
    """`python bench.py` on a tiny workload: one JSON line with every key of the driver's contract, measured through
    the CUDA path (kernel launches counted, the training step replayed from its graph)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "tiny", "--steps", "4",
                        "--warmup", "1", "--no-cpu-baseline", "--eval-batches", "2"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    z = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "eval"):
        assert k in z, k
    assert z["metric"] == "train_seqs_per_sec" and z["unit"] == "seq/s" and z["steps"] == 4 and z["value"] > 0
    assert z["gpu_launches"] > 0 and z["impl"]["cuda_graph_steps"] is True and "workload" in z["config"]
    assert z["impl"]["eval_pad_key_shortcut"] is True and any("K2 SpMM" in e["kernel"] for e in z["roofline_extra"])
    assert set(z["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and z["e2e"]["h2d_bytes_per_step"] > 0
    assert set(z["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert z["eval"]["value"] > 0
