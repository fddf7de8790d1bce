"""Device preprocessor vs the reference's own preprocessing outputs (golden fixtures
written by the reference, dataloader.py:60-228) and vs the host restatement on ragged synthetic logs.  Bit-exact."""
import random

import numpy as np
import pytest
import torch

from helpers import GOLDEN_NAMES, Golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_device_preprocess_matches_reference(name):
    from c2dsr_b200 import dataloader as dl
    g = Golden(name)
    hp = g.hp
    na, nb, L = hp["n_item_a"], hp["n_item_b"], hp["len_max"]
    random.seed(hp["seed"])                      # the reference preprocesses train, val, test in this order
    train = dl.preprocess_train_device(g.raw("train"), na, nb, L, "cuda")
    assert torch.equal(train.cpu(), torch.from_numpy(g.z["train_fields"]))
    for mode in ("val", "test"):
        six, four, neg = dl.preprocess_evaluate_device(g.raw(mode), na, nb, L, hp["n_neg_sample"], "cuda")
        assert torch.equal(six.cpu(), torch.from_numpy(g.z[f"{mode}_six"]))
        assert torch.equal(four.cpu(), torch.from_numpy(g.z[f"{mode}_four"]))
        assert torch.equal(neg.cpu(), torch.from_numpy(g.z[f"{mode}_neg"]))


def _logs(n, na, nb, L, seed):
    r = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        m = int(r.randint(1, L + 2))             # 1 .. L + 1 items: from "a target and no input" to the longest legal
        kind = r.randint(0, 4)
        lo, hi = (0, na + nb) if kind < 2 else ((0, na) if kind == 2 else (na, na + nb))    # mixed / A only / B only
        out.append(r.randint(lo, hi, m).tolist())
    out.append([na] * 3)                         # the first id of domain B as final target: Q16's strict comparison
    out.append([0, na - 1, na, na + nb - 1])
    return out


@pytest.mark.parametrize("na,nb,L,n", [(5, 9, 6, 300), (120, 77, 15, 2000), (40, 41, 50, 500)])
def test_device_preprocess_matches_host_on_ragged_logs(na, nb, L, n):
    from c2dsr_b200 import dataloader as dl
    seqs = _logs(n, na, nb, L, seed=na + L)
    n_neg = 7
    ev = [u for u in seqs if (u[-1] if u[-1] < na else u[-1] - na) + max(0, (na if u[-1] < na else nb - na)
                                                                         - (u[-1] if u[-1] < na else u[-1] - na) - 1)
          >= n_neg]                              # the reference's sample() raises when the population is too small
    rng_h, rng_d = random.Random(7), random.Random(7)
    want = dl.preprocess_train(seqs, na, nb, L, rng=rng_h)
    got = dl.preprocess_train_device(seqs, na, nb, L, "cuda", rng=rng_d)
    assert got.shape == want.shape and 0 < len(want) < len(seqs)
    assert np.array_equal(got.cpu().numpy(), want)
    w6, w4, wn = dl.preprocess_evaluate(ev, na, nb, L, n_neg, rng=rng_h)
    g6, g4, gn = dl.preprocess_evaluate_device(ev, na, nb, L, n_neg, "cuda", rng=rng_d)
    assert np.array_equal(g6.cpu().numpy(), w6) and np.array_equal(g4.cpu().numpy(), w4)
    assert np.array_equal(gn.cpu().numpy(), wn)
    assert rng_h.random() == rng_d.random()      # both consumed the same stream


def test_device_preprocess_edge_cases():
    from c2dsr_b200 import dataloader as dl
    assert dl.preprocess_train_device([], 5, 9, 6, "cuda").shape == (0, 14, 6)
    assert dl.preprocess_train_device([[0, 1, 2, 3]], 5, 9, 6, "cuda").shape[0] == 0      # no domain-B target
    with pytest.raises(ValueError):
        dl.preprocess_train_device([list(range(9))], 5, 9, 6, "cuda")
    six, four, neg = dl.preprocess_evaluate_device([[0, 1, 2]], 5, 12, 6, 3, "cuda")
    assert four.cpu().tolist() == [[5, -1, 0, 2]] and 2 not in neg.cpu().tolist()[0]


def test_main_device_preprocess_matches_host(tmp_path):
    """``--use_raw --device_preprocess`` builds the same splits as the host path (same seed, same ``random`` stream)."""
    from argparse import Namespace
    from c2dsr_b200.dataloader import CDSRDataset
    na, nb, L = 60, 70, 12
    seqs = _logs(400, na, nb, L, seed=3)
    for mode in ("train", "val"):
        with open(tmp_path / f"{mode}_new.txt", "w") as f:
            for u in seqs:
                f.write("0\t0\t" + "\t".join(f"{x}|{t}" for t, x in enumerate(u)) + "\n")
    base = dict(use_raw=True, path_raw=str(tmp_path), path_data=str(tmp_path), n_item_a=na, n_item_b=nb, len_max=L,
                n_neg_sample=9, save_processed=False, device="cuda")
    for mode in ("train", "val"):
        random.seed(11)
        host = CDSRDataset(Namespace(**base), mode)
        random.seed(11)
        dev = CDSRDataset(Namespace(device_preprocess=True, **base), mode)
        assert len(host) == len(dev) and dev.fields[0].is_cuda
        for a, b in zip(host.fields, dev.fields):
            assert torch.equal(a, b.cpu())
