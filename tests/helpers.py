"""Shared test helpers: golden-fixture loading and tolerance helpers."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = ("tiny_default", "tiny_deep", "tiny_prenorm_shared", "mid_default")
sys.path.insert(0, os.path.join(ROOT, "oracle"))


class Golden:
    """One ``tests/golden/<name>.npz`` produced by make_golden.py from the reference."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.hp = json.loads(str(self.z["hp"]))
        self.name = name

    def group(self, prefix):
        return {k[len(prefix) + 1:]: torch.from_numpy(self.z[k]) for k in self.z.files if k.startswith(prefix + "/")}

    def raw(self, mode):
        items, lens = self.z[f"raw_{mode}_items"], self.z[f"raw_{mode}_lens"]
        out, o = [], 0
        for n in lens:
            out.append(items[o:o + n].tolist())
            o += n
        return out

    def adj(self, which):
        n = self.hp["n_item"]
        idx = torch.from_numpy(np.stack((self.z[f"adj_{which}_row"], self.z[f"adj_{which}_col"])))
        return torch.sparse_coo_tensor(idx, torch.from_numpy(self.z[f"adj_{which}_val"]), (n, n)).coalesce()

    def train_batch(self, step):
        f = torch.from_numpy(self.z["train_fields"])
        B = self.hp["batch_size"]
        return tuple(f[step * B:(step + 1) * B, i].contiguous() for i in range(14))

    def eval_batch(self, mode="val"):
        six, four, neg = (torch.from_numpy(self.z[f"{mode}_{x}"]) for x in ("six", "four", "neg"))
        return tuple(six[:, i].contiguous() for i in range(6)) + tuple(four[:, i:i + 1].contiguous()
                                                                       for i in range(4)) + (neg,)


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
