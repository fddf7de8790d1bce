"""Host-side data path vs the reference's own preprocessing outputs (golden fixtures).

Integer work: bit-exact.  Covers dataloader.py:60-228 (train/eval preprocessing with the same
``random`` stream), utils/graph.py:33-96 (adjacency), the DataLoader-compatible shuffle order,
and utils/metrics.py.
"""
import random

import numpy as np
import pytest
import torch

from helpers import GOLDEN_NAMES, Golden
import c2dsr_oracle as oracle
from c2dsr_b200 import dataloader as dl
from c2dsr_b200 import graph as gr
from c2dsr_b200 import metrics, synth


@pytest.fixture(scope="module")
def g():
    return Golden("tiny_default")


def test_preprocess_matches_reference_bit_exact(g):
    hp = g.hp
    random.seed(hp["seed"])                      # make_golden seeds, then Trainer preprocesses train, val, test
    train = dl.preprocess_train(g.raw("train"), hp["n_item_a"], hp["n_item_b"], hp["len_max"])
    assert np.array_equal(train, g.z["train_fields"])
    for mode in ("val", "test"):
        six, four, neg = dl.preprocess_evaluate(g.raw(mode), hp["n_item_a"], hp["n_item_b"], hp["len_max"],
                                                hp["n_neg_sample"])
        assert np.array_equal(six, g.z[f"{mode}_six"])
        assert np.array_equal(four, g.z[f"{mode}_four"])
        assert np.array_equal(neg, g.z[f"{mode}_neg"])


def test_preprocess_edge_cases():
    # empty input, single-domain sequences (dropped from train), over-long sequences (error)
    assert dl.preprocess_train([], 5, 9, 6).shape == (0, 14, 6)
    assert dl.preprocess_train([[0, 1, 2, 3]], 5, 9, 6).shape[0] == 0            # no domain-B target
    with pytest.raises(ValueError):
        dl.preprocess_train([list(range(9))], 5, 9, 6)
    six, four, neg = dl.preprocess_evaluate([[0, 1, 2]], 5, 12, 6, 3)
    assert four.tolist() == [[5, -1, 0, 2]] and six.shape == (1, 6, 6) and 2 not in neg[0]
    with pytest.raises(ValueError):                                              # Q7b: population too small
        dl.preprocess_evaluate([[0, 6]], 5, 7, 6, 3)


def test_graph_matches_reference(g):
    hp = g.hp
    shared, spec = gr.transition_edges(g.raw("train"), hp["n_item_a"])
    for nm, e in (("share", shared), ("spec", spec)):
        r, c, v = gr.normalised_coo(e, hp["n_item"])
        assert np.array_equal(r, g.z[f"adj_{nm}_row"]) and np.array_equal(c, g.z[f"adj_{nm}_col"])
        np.testing.assert_allclose(v, g.z[f"adj_{nm}_val"], rtol=1e-7)
    # and against the oracle's loop restatement on a second random log
    seqs = synth.make_sequences(300, 40, 60, len_max=12, seed=7)
    (a, b) = oracle.build_adjacency(seqs, 40, 101)
    s2, p2 = gr.transition_edges(seqs, 40)
    for ref, e in ((a, s2), (b, p2)):
        r, c, v = gr.normalised_coo(e, 101)
        assert np.array_equal(r, ref[0]) and np.array_equal(c, ref[1])
        np.testing.assert_allclose(v, ref[2], rtol=1e-7)


def test_csr_graph_roundtrip(g):
    adj = g.adj("share")
    cg = gr.CsrGraph(adj, device="cpu")
    dense = adj.to_dense()
    n = dense.shape[0]
    rebuilt = torch.zeros(n, n)
    rows = torch.repeat_interleave(torch.arange(n), (cg.rowptr[1:] - cg.rowptr[:-1]).long())
    rebuilt[rows, cg.col.long()] = cg.val
    assert torch.equal(rebuilt, dense)
    rebuilt_t = torch.zeros(n, n)
    rows = torch.repeat_interleave(torch.arange(n), (cg.t_rowptr[1:] - cg.t_rowptr[:-1]).long())
    rebuilt_t[rows, cg.t_col.long()] = cg.t_val
    assert torch.equal(rebuilt_t, dense.t())


def test_batchloader_order_equals_torch_dataloader():
    class Args:
        len_max = 4
    ds = dl.CDSRDataset.__new__(dl.CDSRDataset)
    ds.fields = [torch.arange(23).view(23, 1), torch.arange(23).view(23, 1) * 2]
    ds.length, ds.mode, ds.len_max = 23, "train", 4
    torch.manual_seed(3407)
    mine = [b[0].view(-1).tolist() for b in dl.BatchLoader(ds, 5, shuffle=True)]
    torch.manual_seed(3407)
    ref = [b[0].view(-1).tolist() for b in torch.utils.data.DataLoader(ds, batch_size=5, shuffle=True)]
    assert mine == ref
    # data-parallel slices partition every global batch
    torch.manual_seed(1)
    whole = [b[0].view(-1).tolist() for b in dl.BatchLoader(ds, 6, shuffle=True)]
    parts = []
    for r in range(2):
        torch.manual_seed(1)
        parts.append([b[0].view(-1).tolist() for b in dl.BatchLoader(ds, 6, shuffle=True, rank=r, world_size=2)])
    assert [a + b for a, b in zip(*parts)] == whole


def test_metrics_equal_oracle():
    rng = np.random.default_rng(0)
    ra, rb = rng.integers(1, 60, 500).tolist(), rng.integers(1, 30, 300).tolist()
    assert metrics.cal_metrics(ra) == oracle.cal_metrics(ra)
    np.testing.assert_allclose(metrics.cal_score(ra, rb, metrics.BENCHMARKS["fk"]),
                               oracle.cal_score(ra, rb, metrics.BENCHMARKS["fk"]), rtol=1e-14)
    with pytest.raises(ZeroDivisionError):
        metrics.cal_metrics([])


def test_device_preprocessor_host_half_follows_the_reference_stream():
    """The draws the device preprocessor takes as input (dataloader.train_draws / eval_picks) are the ones the
    reference's preprocessing consumed: recovered from the golden fields the reference wrote (corrupted sequences,
    negatives) and from the host restatement on ragged synthetic logs, including the stream position afterwards."""
    g = Golden("mid_default")
    hp = g.hp
    na, nb, L = hp["n_item_a"], hp["n_item_b"], hp["len_max"]
    random.seed(hp["seed"])
    seqs = g.raw("train")
    items, offs = dl._flatten(seqs, L)
    draws = dl.train_draws(items, offs, na, nb)
    # sequences the reference kept, in order: its seq_share_neg_a / _b hold the draw at every input position
    fields = g.z["train_fields"]
    kept = [u for u in seqs if len(dl.preprocess_train([u], na, nb, L, rng=random.Random(0)))]
    assert len(kept) == len(fields)
    pos = {id(u): int(offs[i]) - i for i, u in enumerate(seqs)}
    for row, u in zip(fields, kept):
        m = len(u) - 1
        seq = np.asarray(u[:-1])
        want = np.where(seq < na, row[13][L - m:], row[12][L - m:])        # A position: neg_b holds the draw
        assert np.array_equal(draws[pos[id(u)]:pos[id(u)] + m], want)
    for mode in ("val", "test"):                                            # same stream, continued
        ev = g.raw(mode)
        items, offs = dl._flatten(ev, L)
        picks = dl.eval_picks(items, offs, na, nb, hp["n_neg_sample"])
        gt = g.z[f"{mode}_four"][:, 3:4]
        assert np.array_equal(np.where(picks < gt, picks, picks + 1), g.z[f"{mode}_neg"])
    # ragged synthetic log: same consumption as the host restatement
    r = np.random.RandomState(5)
    logs = [r.randint(0, 40, int(r.randint(1, 9))).tolist() for _ in range(200)]
    a, b = random.Random(3), random.Random(3)
    dl.preprocess_train(logs, 17, 23, 7, rng=a)
    items, offs = dl._flatten(logs, 7)
    assert len(dl.train_draws(items, offs, 17, 23, rng=b)) == len(items) - len(logs)
    assert a.random() == b.random()
    with pytest.raises(ValueError):
        dl._flatten([[1] * 10], 7)
