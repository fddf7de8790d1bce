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
