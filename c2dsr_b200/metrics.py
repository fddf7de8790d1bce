"""Recall(HR) / MRR / NDCG @ {5, 20} from integer ranks (reference: utils/metrics.py)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

MAPPING_DATASET = {"fk": "Food-Kitchen", "mb": "Movie-Book", "ee": "Entertainment-Education"}
# paper numbers [HR@5_A, NDCG@5_A, HR@5_B, NDCG@5_B] the reference divides by (utils/constant.py:13-17)
BENCHMARKS = {"fk": [0.1124, 0.0865, 0.0574, 0.0416], "mb": [0.0647, 0.0476, 0.0284, 0.0217],
              "ee": [0.6945, 0.5620, 0.7104, 0.5905]}


def cal_metrics(ranks: Sequence[int]) -> List[float]:
    """[hr5, hr20, mrr5, mrr20, ndcg5, ndcg20]; means of 1[r<=k], 1[r<=k]/r, 1[r<=k]/log2(r+1)
    (utils/metrics.py:4-19).  An empty list raises ZeroDivisionError like the reference."""
    n = len(ranks)
    if n == 0:
        raise ZeroDivisionError("cal_metrics on an empty rank list")
    r = np.asarray(ranks, dtype=np.int64)
    out = []
    for f in (lambda x: np.ones_like(x, dtype=np.float64), lambda x: 1.0 / x, lambda x: 1.0 / np.log2(x + 1.0)):
        for k in (5, 20):
            hit = r[r <= k]
            out.append(float(np.cumsum(f(hit.astype(np.float64)))[-1] / n) if hit.size else 0.0)
    return out


def cal_score(ranks_a, ranks_b, benchmark) -> List[float]:
    """[mean improvement over the paper's four numbers] + 6 metrics of A + 6 of B (utils/metrics.py:22-31)."""
    res = cal_metrics(ranks_a) + cal_metrics(ranks_b)
    picked = np.array([res[0], res[4], res[6], res[10]])
    return [float(np.mean(picked / np.asarray(benchmark, dtype=np.float64) - 1))] + res
