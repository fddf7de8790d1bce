"""Generate the golden fixtures by RUNNING THE REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only checkout)

The reference (crystal22/C2DSR) ships no tests or golden vectors, so the parity pin is
produced here: tiny seeded synthetic logs in the reference's raw text format are pushed
through the reference's own ``CDSRDataset`` / ``preprocess_graph`` / ``C2DSR`` /
``Trainer`` (CPU, dropouts 0, torch 2.11.0), and inputs + outputs are stored as
``tests/golden/<name>.npz``.  Nothing at test time imports the reference; the tests
read only the ``.npz`` files.  Reference entry points exercised:
  dataloader.py:60-228, utils/graph.py:33-96, models/C2DSR.py:9-85,
  trainer.py:40-181, utils/metrics.py:4-31.
"""
import argparse
import json
import os
import random
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("C2DSR_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

CONFIGS = {
    # name: overrides of the reference's default hyper-parameters (main.py:18-65)
    "tiny_default": dict(),
    "tiny_deep": dict(n_gnn=2, n_attn=2, n_head=2, d_bias=True),
    "tiny_prenorm_shared": dict(n_attn=2, n_head=4, norm_first=True, shared_item_embed=True),
    # more steps (eager, eager, capture, replays of Trainer.train_step) and a validation split large enough for
    # the 1e-3 metric bar to mean something (> 1 000 queries per domain); keys starting with "_" size the data
    "mid_default": dict(len_rec=10, len_max=15, n_neg_sample=99, _na=300, _nb=450, _train=300, _val=2400, _test=40,
                        _steps=8, _eval_acts=False),
}


def build_args(root, over):
    a = argparse.Namespace(
        data="fk", dataset="Food-Kitchen", len_rec=4, use_raw=True, save_processed=True, n_neg_sample=12,
        zip_ee=False, d_latent=32, disable_embed_l2=False, shared_item_embed=False, d_bias=False, n_gnn=1,
        dropout_gnn=0.0, n_attn=1, n_head=1, dropout_attn=0.0, norm_first=False, lr=1e-3, lr_decay=0.1, l2=5e-4,
        lr_gamma=0.5, lr_step=10, n_lr_decay=5, decay_epoch=5, max_grad_norm=5.0, len_max=10, lambda_loss=0.7,
        cuda="cpu", seed=3407, n_epoch=1, batch_size=32, batch_size_eval=64, num_workers=0, es_patience=10,
        device=torch.device("cpu"), path_root=root, path_data=os.path.join(root, "data", "Food-Kitchen"),
        path_raw=os.path.join(root, "data", "raw", "Food-Kitchen"), path_ckpt=os.path.join(root, "checkpoints"),
        path_log=os.path.join(root, "log"), benchmark=[0.1124, 0.0865, 0.0574, 0.0416])
    for k, v in over.items():
        if not k.startswith("_"):
            setattr(a, k, v)
    for p in (a.path_data, a.path_raw, a.path_ckpt, a.path_log):
        os.makedirs(p, exist_ok=True)
    return a


def ragged(seqs):
    return np.concatenate([np.asarray(s, np.int64) for s in seqs]), np.asarray([len(s) for s in seqs], np.int64)


def coo_of(t):
    t = t.coalesce()
    return t.indices()[0].numpy(), t.indices()[1].numpy(), t.values().numpy()


def make_one(name, over):
    from c2dsr_b200 import synth
    sys.path.insert(0, REF)
    from dataloader import CDSRDataset, get_dataloader          # noqa: reference modules
    from trainer import Trainer
    from utils.metrics import cal_score

    class Quiet:                                               # stands in for utils/noter.py Noter
        def log_train(self, *a):
            pass

    NA, NB = over.get("_na", 50), over.get("_nb", 71)
    n_steps = over.get("_steps", 3)
    root = tempfile.mkdtemp(prefix="c2dsr_golden_")
    args = build_args(root, over)
    raw = {"train": synth.make_sequences(over.get("_train", 150), NA, NB, len_max=args.len_max, seed=1),
           "val": synth.make_sequences(over.get("_val", 48), NA, NB, len_max=args.len_max, seed=2),
           "test": synth.make_sequences(over.get("_test", 40), NA, NB, len_max=args.len_max, seed=3)}
    raw["train"][0][-1] = NA                                   # exercise the id == n_item_a corner (Q16)
    for mode, seqs in raw.items():
        synth.write_raw(os.path.join(args.path_raw, mode + "_new.txt"), seqs)
    synth.write_item_lists(args.path_raw, NA, NB)

    random.seed(args.seed); torch.manual_seed(args.seed); np.random.seed(args.seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tr = Trainer(args, Quiet())                            # runs the reference preprocessors + model init
    out = {"hp": json.dumps({k: v for k, v in vars(args).items() if isinstance(v, (int, float, bool, str))}
                            | dict(n_item_a=args.n_item_a, n_item_b=args.n_item_b, n_item=args.n_item,
                                   idx_pad=args.idx_pad))}
    for mode, seqs in raw.items():
        out[f"raw_{mode}_items"], out[f"raw_{mode}_lens"] = ragged(seqs)
    ds = {"train": tr.trainloader.dataset, "val": tr.valloader.dataset, "test": tr.testloader.dataset}
    out["train_fields"] = np.asarray(ds["train"].data, np.int64)                      # [n, 14, L]
    for mode in ("val", "test"):
        d = ds[mode].data
        out[f"{mode}_six"] = np.asarray([r[:6] for r in d], np.int64)
        out[f"{mode}_four"] = np.asarray([[r[6][0], r[7][0], r[8][0], r[9][0]] for r in d], np.int64)
        out[f"{mode}_neg"] = np.asarray([r[10] for r in d], np.int64)
    for nm, adj in (("share", tr.adj_share), ("spec", tr.adj_specific)):
        out[f"adj_{nm}_row"], out[f"adj_{nm}_col"], out[f"adj_{nm}_val"] = coo_of(adj)

    sd0 = {k: v.detach().clone() for k, v in tr.model.state_dict().items()}
    for k, v in sd0.items():
        out["init/" + k] = v.numpy()

    # --- K training steps, fixed batch order (contiguous slices), trainer.py:40-53 ---------
    fields = torch.from_numpy(out["train_fields"])
    B = args.batch_size
    tr.model.train()
    tr.optimizer.zero_grad()
    losses = []
    for s in range(n_steps):
        batch = tuple(fields[s * B:(s + 1) * B, i] for i in range(14))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr.model.convolve_graph()
            if s == 0:
                with torch.no_grad():
                    hs, hx, hy = tr.model(*batch[:6])
                    out["step0/h_share"], out["step0/hx"], out["step0/hy"] = hs.numpy(), hx.numpy(), hy.numpy()
                    out["step0/hi_share"] = tr.model.hi_share.detach().numpy()
                    out["step0/hi_a"] = tr.model.hi_a.detach().numpy()
                    out["step0/hi_b"] = tr.model.hi_b.detach().numpy()
                    out["step0/h_neg_a"] = tr.model.forward_share(batch[12], batch[3]).numpy()
            loss = tr.train_batch(batch)
        losses.append([float(x) for x in loss])
        if s == 0:
            for k, p in tr.model.named_parameters():
                if p.grad is not None:
                    out["grad0/" + k] = p.grad.detach().clone().numpy()
    out["losses"] = np.asarray(losses, np.float64)
    for k, v in tr.model.state_dict().items():
        out["final/" + k] = v.detach().numpy()

    # --- evaluation on the val split, trainer.py:61-70,162-181 -----------------------------
    tr.model.eval()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tr.model.convolve_graph()
        six, four, neg = (torch.from_numpy(out[f"val_{x}"]) for x in ("six", "four", "neg"))
        batch = tuple(six[:, i] for i in range(6)) + tuple(four[:, i:i + 1] for i in range(4)) + (neg,)
        ra, rb = tr.evaluate_batch(batch)
        if over.get("_eval_acts", True):
            hs, hx, hy = tr.model(*batch[:6])
            out["eval/h_share"], out["eval/hx"], out["eval/hy"] = hs.numpy(), hx.numpy(), hy.numpy()
    out["eval/rank_a"], out["eval/rank_b"] = np.asarray(ra, np.int64), np.asarray(rb, np.int64)
    out["eval/score"] = np.asarray(cal_score(ra, rb, args.benchmark), np.float64)

    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: losses {losses}  ranks_a {len(ra)} ranks_b {len(rb)}  -> {path} "
          f"({os.path.getsize(path) / 1e3:.0f} kB)")
    for m in [m for m in list(sys.modules) if m.split(".")[0] in ("trainer", "dataloader", "models", "utils")]:
        del sys.modules[m]


if __name__ == "__main__":
    only = sys.argv[1:]
    for nm, ov in CONFIGS.items():
        if not only or nm in only:
            make_one(nm, ov)
