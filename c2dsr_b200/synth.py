"""Synthetic cross-domain interaction logs shaped like the reference's datasets.

The reference's real training files are absent, so benchmarks and
parity runs use seeded synthetic logs with the published catalogue sizes, the length
histogram of the shipped ``val_new.txt`` and a Zipf item popularity.
Logs can be written in the reference's raw text format (Appendix B) so that they can
be pushed through either preprocessor.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

# catalogue sizes and split sizes: img/data.png, data/raw/leak_stats.py:123-153 of the reference
SHAPES = {
    "fk": dict(dataset="Food-Kitchen", n_item_a=29207, n_item_b=34886, n_train=34117, n_val=8173, n_test=8406,
               len_max=15, frac_a=0.48),
    "mb": dict(dataset="Movie-Book", n_item_a=36845, n_item_b=63937, n_train=58515, n_val=7644, n_test=7708,
               len_max=15, frac_a=0.50),
    "ee": dict(dataset="Entertainment-Education", n_item_a=8367, n_item_b=11404, n_train=120635, n_val=6929,
               n_test=6785, len_max=30, frac_a=0.50),
    "tiny": dict(dataset="Synthetic-tiny", n_item_a=1500, n_item_b=2500, n_train=600, n_val=64, n_test=64,
                 len_max=15, frac_a=0.48),
    "1m": dict(dataset="Synthetic-1M", n_item_a=400000, n_item_b=600000, n_train=0, n_val=8192, n_test=8192,
               len_max=15, frac_a=0.45),
}
# sequence length (incl. target) histogram of Food-Kitchen val_new.txt, lengths 6..15
FK_LEN_HIST = {6: 1130, 7: 1215, 8: 1137, 9: 918, 10: 717, 11: 629, 12: 509, 13: 389, 14: 385, 15: 1144}


def _zipf_sampler(n: int, alpha: float, rng: np.random.Generator, offset: float = 50.0):
    # Zipf-Mandelbrot: p(r) ~ (r + offset)^-alpha.  The offset flattens the head so that the most popular
    # item is seen ~10^2 times at Food-Kitchen size, like the shipped val_new.txt (max out-degree ~10^2,
    # SURVEY.md section 8(d)); a pure power law would give a 10x heavier head than the real logs.
    p = (np.arange(1, n + 1, dtype=np.float64) + offset) ** (-alpha)
    cdf = np.cumsum(p / p.sum())
    perm = rng.permutation(n)                       # popularity is not correlated with the id
    return lambda size: perm[np.minimum(np.searchsorted(cdf, rng.random(size)), n - 1)]


def make_sequences(n_seq: int, n_item_a: int, n_item_b: int, len_max: int = 15, frac_a: float = 0.48,
                   alpha: float = 0.65, seed: int = 0, lengths: str = "fk") -> List[List[int]]:
    """``n_seq`` time-ordered item lists (global ids; A items < n_item_a <= B items).

    lengths: "fk" = the Food-Kitchen histogram scaled to ``len_max``; "full" = all
    sequences have the maximum length ``len_max + 1`` (Entertainment-Education style);
    "uniform" = U[6, len_max + 1] (the hidden-size / length sweep).
    """
    rng = np.random.default_rng(seed)
    top = len_max + 1 if lengths != "fk" else len_max
    if lengths == "fk":
        ls = np.array(sorted(FK_LEN_HIST))
        pr = np.array([FK_LEN_HIST[k] for k in ls], dtype=np.float64)
        lens = rng.choice(ls, size=n_seq, p=pr / pr.sum())
        if len_max != 15:
            lens = np.clip(np.round(lens * (len_max / 15.0)).astype(int), 4, len_max)
    elif lengths == "full":
        lens = np.full(n_seq, top)
    else:
        lens = rng.integers(6, top + 1, size=n_seq)
    total = int(lens.sum())
    pick_a, pick_b = _zipf_sampler(n_item_a, alpha, rng), _zipf_sampler(n_item_b, alpha, rng)
    is_a = rng.random(total) < frac_a
    items = np.where(is_a, pick_a(total), pick_b(total) + n_item_a)
    out, o = [], 0
    for n in lens:
        out.append(items[o:o + n].tolist())
        o += n
    return out


def write_raw(path: str, seqs: Sequence[Sequence[int]], t0: int = 1_300_000_000) -> None:
    """Raw text of Appendix B: ``user \\t inter_id \\t item|unix_ts|date| \\t ...``."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w", encoding="utf-8") as f:
        for u, seq in enumerate(seqs):
            toks = [f"{it}|{t0 + 86400 * (u % 1000) + 60 * k}|1970-01-01 00:00:00|" for k, it in enumerate(seq)]
            f.write(f"{u}\t{u}\t" + "\t".join(toks) + "\n")


def write_item_lists(dirpath: str, n_item_a: int, n_item_b: int) -> None:
    """``items_a.txt`` / ``items_b.txt``: only the line count is ever used."""
    os.makedirs(dirpath, exist_ok=True)
    for name, n in (("items_a.txt", n_item_a), ("items_b.txt", n_item_b)):
        with open(os.path.join(dirpath, name), "w", encoding="utf-8") as f:
            f.writelines(f"{i}\tX{i}\t{i}\n" for i in range(n))
