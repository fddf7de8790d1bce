"""Run the UNMODIFIED reference from ``oracle/_ref`` on the host CPU  --  TEST / BENCH INFRASTRUCTURE.

Used by ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg (``kind: "reference"``) and by tests/.
The reference's own classes are driven exactly as its ``main.py`` / ``Trainer.run_epoch`` drive them
(trainer.py:40-71): ``Trainer(args, noter)`` built from processed pickles in the reference's on-disk layout
(dataloader.py:29-35, utils/graph.py:103-107), then per step ``model.convolve_graph()`` + ``train_batch(batch)``.
Nothing of this repository's package is on that path; the synthetic pickles are written by the caller.
"""
from __future__ import annotations

import argparse
import os
import pickle
import sys
import time
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "trainer.py"))


class _Quiet:                                           # stands in for utils/noter.py Noter
    def log_train(self, *a):
        pass

    def log_msg(self, *a):
        pass


def write_processed(root: str, dataset: str, n_item_a: int, n_item_b: int, train_fields: np.ndarray, ev, adj):
    """``data/<dataset>/{items_a.txt, items_b.txt, train.pkl, val.pkl, test.pkl, graph.pkl}`` in the reference's
    formats: a split is a list of samples, each a list of per-field int lists (dataloader.py:159-160, 218-226)."""
    p = os.path.join(root, "data", dataset)
    os.makedirs(p, exist_ok=True)
    for name, n in (("items_a.txt", n_item_a), ("items_b.txt", n_item_b)):
        with open(os.path.join(p, name), "w", encoding="utf-8") as f:
            f.writelines(f"{i}\tX{i}\t{i}\n" for i in range(n))
    with open(os.path.join(p, "train.pkl"), "wb") as f:
        pickle.dump(train_fields.tolist(), f)
    six, four, neg = ev
    rows = [[*six[i].tolist(), *[[int(v)] for v in four[i]], neg[i].tolist()] for i in range(len(six))]
    for mode in ("val", "test"):
        with open(os.path.join(p, mode + ".pkl"), "wb") as f:
            pickle.dump(rows, f)
    with open(os.path.join(p, "graph.pkl"), "wb") as f:
        pickle.dump((adj[0].cpu(), adj[1].cpu()), f)
    return p


def reference_args(root: str, hp: dict) -> argparse.Namespace:
    """The Namespace main.py:69-89 derives, for a processed (not --use_raw) CPU run."""
    a = argparse.Namespace(
        data=hp["data"], dataset=hp["dataset"], len_rec=hp["len_rec"], use_raw=False, save_processed=False,
        n_neg_sample=hp["n_neg_sample"], zip_ee=False, d_latent=hp["d_latent"], disable_embed_l2=False,
        shared_item_embed=hp.get("shared_item_embed", False), d_bias=hp.get("d_bias", False), n_gnn=hp["n_gnn"],
        dropout_gnn=hp["dropout_gnn"], n_attn=hp["n_attn"], n_head=hp["n_head"], dropout_attn=hp["dropout_attn"],
        norm_first=hp.get("norm_first", False), lr=hp["lr"], lr_decay=0.1, l2=hp["l2"], lr_gamma=hp["lr_gamma"],
        lr_step=hp["lr_step"], n_lr_decay=5, decay_epoch=5, max_grad_norm=5.0, len_max=hp["len_max"],
        lambda_loss=hp["lambda_loss"], cuda="cpu", seed=hp["seed"], n_epoch=1, batch_size=hp["batch_size"],
        batch_size_eval=hp["batch_size_eval"], num_workers=0, es_patience=10, device=torch.device("cpu"),
        path_root=root, path_data=os.path.join(root, "data", hp["dataset"]),
        path_raw=os.path.join(root, "data", "raw", hp["dataset"]), path_ckpt=os.path.join(root, "checkpoints"),
        path_log=os.path.join(root, "log"))
    return a


def load_trainer(root: str, hp: dict):
    """Import the reference from oracle/_ref and build its Trainer (trainer.py:13-38) on CPU."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    args = reference_args(root, hp)
    torch.manual_seed(args.seed)
    import contextlib
    with warnings.catch_warnings(), contextlib.redirect_stdout(sys.stderr):      # (the reference prints progress lines)
        warnings.simplefilter("ignore")
        import trainer as ref_trainer                   # noqa: the reference's module, from oracle/_ref
        tr = ref_trainer.Trainer(args, _Quiet())
    return tr, args


def time_reference(root: str, hp: dict, train_fields: np.ndarray, steps: int, warmup: int, eval_queries: int = 128,
                   max_seconds: float | None = None):
    """-> dict(train seq/s, eval q/s, steps actually run).  A step = the body of trainer.py:47-49."""
    torch.set_num_threads(os.cpu_count() or 1)
    tr, args = load_trainer(root, hp)
    B = args.batch_size
    n_b = max(1, train_fields.shape[0] // B)
    batches = [tuple(torch.from_numpy(np.ascontiguousarray(train_fields[i * B:(i + 1) * B, f])) for f in range(14))
               for i in range(min(n_b, 4))]
    tr.model.train()
    tr.optimizer.zero_grad()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(warmup):
            tr.model.convolve_graph()
            tr.train_batch(batches[i % len(batches)])
        t0, n, loss = time.perf_counter(), 0, (float("nan"),)
        while n < steps:
            tr.model.convolve_graph()
            loss = tr.train_batch(batches[n % len(batches)])
            float(loss[0])
            n += 1
            if max_seconds is not None and time.perf_counter() - t0 > max_seconds:
                break
        el = max(time.perf_counter() - t0, 1e-9)
        # evaluation: trainer.py:61-70 on a bounded slice of the validation split (the reference computes the
        # full-domain score vector per query and counts over list_neg: a lower bound for full-catalogue ranking)
        tr.model.eval()
        ds = tr.valloader.dataset
        nq = min(eval_queries, len(ds))
        eb = tuple(torch.stack(x) for x in zip(*[ds[i] for i in range(nq)]))
        with torch.no_grad():
            tr.model.convolve_graph()
            t1 = time.perf_counter()
            ra, rb = tr.evaluate_batch(eb)
            ev_el = time.perf_counter() - t1
    return dict(train_seq_per_s=n * B / el, seconds=el, steps=n, warmup=warmup, batch=B,
                eval_q_per_s=nq / ev_el, eval_queries=nq, threads=torch.get_num_threads(), loss=float(loss[0]))
