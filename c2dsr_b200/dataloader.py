"""Batch pipeline for the C2DSR hot path (host side of SURVEY.md section 8 rows a0 and f-1/f-4).

Mirrors the reference's ``dataloader.py`` interface -- ``CDSRDataset(args, mode)``,
``get_dataloader(args)`` -- and produces *identical* field values (same consumption of
Python's ``random`` stream, so negatives and corruptions match the reference bit for
bit), but keeps every split as one dense int64 tensor ``[n, L]`` per field instead
of a list of per-sample Python lists.  A ``BatchLoader`` then slices whole batches
out of those tensors (optionally already resident on the GPU), replacing the
reference's 14 ``LongTensor(list)`` constructions per sample.

Field order (the hot path's input contract):
  train  (dataloader.py:159-160): seq_share, seq_share_a, seq_share_b, pos, pos_a, pos_b,
         gt_share_a, gt_share_b, gt_a, gt_b, gt_mask_a, gt_mask_b, seq_share_neg_a, seq_share_neg_b
  eval   (dataloader.py:218-226): the first six, then idx_last_a, idx_last_b, xory_last, gt_last [n,1],
         list_neg [n, n_neg]
"""
from __future__ import annotations

import pickle
import random
from os.path import join
from typing import List, Sequence, Tuple

import numpy as np
import torch

TRAIN_FIELDS = ("seq_share", "seq_share_a", "seq_share_b", "pos", "pos_a", "pos_b", "gt_share_a", "gt_share_b",
                "gt_a", "gt_b", "gt_mask_a", "gt_mask_b", "seq_share_neg_a", "seq_share_neg_b")
EVAL_FIELDS = ("seq_share", "seq_share_a", "seq_share_b", "pos", "pos_a", "pos_b", "idx_last_a", "idx_last_b",
               "xory_last", "gt_last", "list_neg")


def read_raw(path: str) -> List[List[int]]:
    """Item lists sorted by timestamp with a stable sort (dataloader.py:39-58)."""
    out = []
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            pairs = [tok.split("|")[:2] for tok in line.strip().split("\t")[2:]]
            items = np.fromiter((int(p[0]) for p in pairs), np.int64, len(pairs))
            ts = np.fromiter((int(p[1]) for p in pairs), np.int64, len(pairs))
            out.append(items[np.argsort(ts, kind="stable")].tolist())
    return out


def _split_domains(seq: np.ndarray, n_item_a: int, pad: int):
    """Per-domain views of a mixed sequence: items kept in place, the rest PAD; 1-based positions."""
    is_a = seq < n_item_a
    seq_a, seq_b = np.where(is_a, seq, pad), np.where(is_a, pad, seq)
    pos_a, pos_b = np.where(is_a, np.cumsum(is_a), 0), np.where(is_a, 0, np.cumsum(~is_a))
    return is_a, seq_a, seq_b, pos_a, pos_b


def _left_pad(x: np.ndarray, L: int, fill: int) -> np.ndarray:
    out = np.full(L, fill, np.int64)
    if len(x):
        out[L - len(x):] = x
    return out


def preprocess_train(seqs: Sequence[Sequence[int]], n_item_a: int, n_item_b: int, len_max: int,
                     rng=random) -> np.ndarray:
    """[n_kept, 14, len_max] int64, values equal to dataloader.py:60-161 of the reference."""
    pad = n_item_a + n_item_b
    rows = []
    for u in seqs:
        if len(u) - 1 > len_max:
            raise ValueError(f"sequence of {len(u)} items exceeds len_max+1 = {len_max + 1}")
        u = np.asarray(u, np.int64)
        seq, gt = u[:-1], u[1:]
        m = len(seq)
        is_a, seq_a, seq_b, pos_a, pos_b = _split_domains(seq, n_item_a, pad)
        # corrupted sequences: one draw per position, in order (dataloader.py:80,85)
        neg_a, neg_b = seq.copy(), seq.copy()
        for i in range(m):
            if is_a[i]:
                neg_b[i] = rng.randint(0, n_item_a - 1)
            else:
                neg_a[i] = rng.randint(n_item_a, pad - 1)
        gt_a, gt_b = np.full(m, n_item_a, np.int64), np.full(m, n_item_b, np.int64)
        ia, ib = np.flatnonzero(is_a), np.flatnonzero(~is_a)
        # next same-domain item is the step target; the last one takes the final target if it
        # belongs to the domain, else it is blanked from the input (dataloader.py:96-130, Q16)
        if len(ia):
            gt_a[ia[:-1]] = seq[ia[1:]]
            if gt[-1] < n_item_a:
                gt_a[ia[-1]] = gt[-1]
            else:
                seq_a[ia[-1]], pos_a[ia[-1]] = pad, 0
        mask_a = (gt_a != n_item_a).astype(np.int64)
        if mask_a.sum() == 0:
            continue
        if len(ib):
            gt_b[ib[:-1]] = seq[ib[1:]] - n_item_a
            if gt[-1] > n_item_a:
                gt_b[ib[-1]] = gt[-1] - n_item_a
            else:
                seq_b[ib[-1]], pos_b[ib[-1]] = pad, 0
        mask_b = (gt_b != n_item_b).astype(np.int64)
        if mask_b.sum() == 0:
            continue
        gt_share_a = np.where(gt < n_item_a, gt, n_item_a)
        gt_share_b = np.where(gt >= n_item_a, gt - n_item_a, n_item_b)
        L = len_max
        rows.append(np.stack([
            _left_pad(seq, L, pad), _left_pad(seq_a, L, pad), _left_pad(seq_b, L, pad),
            _left_pad(np.arange(1, m + 1), L, 0), _left_pad(pos_a, L, 0), _left_pad(pos_b, L, 0),
            _left_pad(gt_share_a, L, n_item_a), _left_pad(gt_share_b, L, n_item_b),
            _left_pad(gt_a, L, n_item_a), _left_pad(gt_b, L, n_item_b),
            _left_pad(mask_a, L, 0), _left_pad(mask_b, L, 0),
            _left_pad(neg_a, L, pad), _left_pad(neg_b, L, pad)]))
    return np.stack(rows) if rows else np.zeros((0, 14, len_max), np.int64)


def preprocess_evaluate(seqs: Sequence[Sequence[int]], n_item_a: int, n_item_b: int, len_max: int,
                        n_neg_sample: int, rng=random) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(seqs [n, 6, L], scalars [n, 4], list_neg [n, n_neg]) equal to dataloader.py:163-228."""
    pad = n_item_a + n_item_b
    six, four, negs = [], [], []
    for u in seqs:
        if len(u) - 1 > len_max:
            raise ValueError(f"sequence of {len(u)} items exceeds len_max+1 = {len_max + 1}")
        u = np.asarray(u, np.int64)
        seq, gt_last = u[:-1], int(u[-1])
        m, L = len(seq), len_max
        is_a, seq_a, seq_b, pos_a, pos_b = _split_domains(seq, n_item_a, pad)
        ia, ib = np.flatnonzero(is_a), np.flatnonzero(~is_a)
        idx_last_a = int(ia[-1]) + (L - m) if len(ia) else -1
        idx_last_b = int(ib[-1]) + (L - m) if len(ib) else -1
        # negatives: a sample of the domain's ids without the target.  Domain B draws from
        # range(n_item_b - n_item_a) (Q7b).  Sampling positions of an index range and shifting
        # past the target is the same draw as sampling the reference's concatenated list.
        if gt_last < n_item_a:
            g, hi, dom = gt_last, n_item_a, 0
        else:
            g, hi, dom = gt_last - n_item_a, n_item_b - n_item_a, 1
        n_pop = g + max(0, hi - g - 1)
        picks = np.asarray(rng.sample(range(n_pop), n_neg_sample), np.int64)
        negs.append(np.where(picks < g, picks, picks + 1))
        six.append(np.stack([_left_pad(seq, L, pad), _left_pad(seq_a, L, pad), _left_pad(seq_b, L, pad),
                             _left_pad(np.arange(1, m + 1), L, 0), _left_pad(pos_a, L, 0), _left_pad(pos_b, L, 0)]))
        four.append([idx_last_a, idx_last_b, dom, g])
    n = len(six)
    return (np.stack(six) if n else np.zeros((0, 6, len_max), np.int64),
            np.asarray(four, np.int64).reshape(n, 4),
            np.stack(negs) if n else np.zeros((0, n_neg_sample), np.int64))


def _flatten(seqs: Sequence[Sequence[int]], len_max: int):
    lens = np.fromiter((len(u) for u in seqs), np.int64, len(seqs))
    if len(lens) and int(lens.max()) - 1 > len_max:
        raise ValueError(f"sequence of {int(lens.max())} items exceeds len_max+1 = {len_max + 1}")
    if len(lens) and int(lens.min()) < 1:
        raise ValueError("empty sequence")
    offs = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    items = np.fromiter((x for u in seqs for x in u), np.int64, int(offs[-1]))
    return items, offs


def train_draws(items: np.ndarray, offs: np.ndarray, n_item_a: int, n_item_b: int, rng=random) -> np.ndarray:
    """The corruption draws of the training split, one per input position in sequence order (the final target of a
    sequence gets none): ``randint(0, n_item_a - 1)`` where the position holds a domain-A item, else
    ``randint(n_item_a, n_item_a + n_item_b - 1)`` (dataloader.py:80,85).  The reference draws for every sequence,
    kept or dropped, so the stream position after a split matches too."""
    pad = n_item_a + n_item_b
    is_input = np.ones(len(items), bool)
    if len(offs) > 1:
        is_input[offs[1:] - 1] = False
    dom_a = (items[is_input] < n_item_a).tolist()
    randint, a_hi, b_lo, b_hi = rng.randint, n_item_a - 1, n_item_a, pad - 1
    return np.fromiter((randint(0, a_hi) if a else randint(b_lo, b_hi) for a in dom_a), np.int64, len(dom_a))


def eval_picks(items: np.ndarray, offs: np.ndarray, n_item_a: int, n_item_b: int, n_neg_sample: int,
               rng=random) -> np.ndarray:
    """[n, n_neg] samples of range(population) per evaluation sequence (dataloader.py:216-224): the population is the
    target's domain without the target (domain B: range(n_item_b - n_item_a), Q7b); shifting the picks past the
    target (done on the device) gives the reference's negatives."""
    n = len(offs) - 1
    last = items[offs[1:] - 1] if n else np.zeros(0, np.int64)
    g = np.where(last < n_item_a, last, last - n_item_a)
    hi = np.where(last < n_item_a, n_item_a, n_item_b - n_item_a)
    n_pop = (g + np.maximum(0, hi - g - 1)).tolist()
    sample = rng.sample
    return np.asarray([sample(range(p), n_neg_sample) for p in n_pop], np.int64).reshape(n, n_neg_sample)


def preprocess_train_device(seqs: Sequence[Sequence[int]], n_item_a: int, n_item_b: int, len_max: int, device,
                            rng=random) -> torch.Tensor:
    """``preprocess_train`` with the per-sequence work on the GPU ([n_kept, 14, len_max] int64 on ``device``).  The
    host keeps only what has to follow Python's ``random`` stream (``train_draws``)."""
    from . import ops
    items, offs = _flatten(seqs, len_max)
    draws = train_draws(items, offs, n_item_a, n_item_b, rng)
    dev = torch.device(device)
    fields, keep = ops.preprocess_train(torch.from_numpy(items).to(dev), torch.from_numpy(offs).to(dev),
                                        torch.from_numpy(draws).to(dev), n_item_a, n_item_b, len_max)
    return fields[keep]


def preprocess_evaluate_device(seqs: Sequence[Sequence[int]], n_item_a: int, n_item_b: int, len_max: int,
                               n_neg_sample: int, device, rng=random):
    """``preprocess_evaluate`` with the per-sequence work on the GPU; the host draws ``rng.sample`` per sequence
    (``eval_picks``) and nothing else."""
    from . import ops
    items, offs = _flatten(seqs, len_max)
    picks = eval_picks(items, offs, n_item_a, n_item_b, n_neg_sample, rng)
    dev = torch.device(device)
    return ops.preprocess_eval(torch.from_numpy(items).to(dev), torch.from_numpy(offs).to(dev),
                               torch.from_numpy(picks).to(dev), n_item_a, n_item_b, len_max)


class CDSRDataset(torch.utils.data.Dataset):
    """Same constructor and item contract as the reference's ``CDSRDataset`` (dataloader.py:9-37,
    230-234); holds one int64 tensor per field instead of nested lists (``self.fields``)."""

    def __init__(self, args, mode: str):
        self.mode = mode
        self.len_max = args.len_max
        if getattr(args, "use_raw", False):
            seqs = read_raw(join(args.path_raw, mode + "_new.txt"))
            dev = getattr(args, "device", None) if getattr(args, "device_preprocess", False) else None
            if mode == "train":
                packed = preprocess_train(seqs, args.n_item_a, args.n_item_b, args.len_max) if dev is None else \
                    preprocess_train_device(seqs, args.n_item_a, args.n_item_b, args.len_max, dev)
                fields = [packed[:, i] for i in range(14)]
            else:
                six, four, neg = preprocess_evaluate(seqs, args.n_item_a, args.n_item_b, args.len_max,
                                                     args.n_neg_sample) if dev is None else \
                    preprocess_evaluate_device(seqs, args.n_item_a, args.n_item_b, args.len_max, args.n_neg_sample,
                                               dev)
                fields = [six[:, i] for i in range(6)] + [four[:, i:i + 1] for i in range(4)] + [neg]
            if getattr(args, "save_processed", True):
                with open(join(args.path_data, mode + ".pkl"), "wb") as f:   # the reference's pickle layout
                    pickle.dump([[fld[i].tolist() for fld in fields] for i in range(len(fields[0]))], f)
        else:
            with open(join(args.path_data, mode + ".pkl"), "rb") as f:
                data = pickle.load(f)
            n_f = 14 if mode == "train" else 11
            fields = [np.asarray([row[i] for row in data], np.int64).reshape(len(data), -1) for i in range(n_f)]
        self.fields = [x.contiguous() if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))
                       for x in fields]
        self.length = len(self.fields[0])

    @classmethod
    def from_fields(cls, fields, mode: str, len_max: int):
        """Wrap already-preprocessed field tensors (train: 14 x [n, L]; eval: 6 x [n, L], 4 x [n, 1], [n, n_neg])."""
        self = cls.__new__(cls)
        self.mode, self.len_max = mode, len_max
        self.fields = [torch.as_tensor(x).contiguous() for x in fields]
        self.length = len(self.fields[0])
        return self

    def valid_counts(self, len_rec: int, na: int, nb: int):
        """Host-side int64 [n, 2]: per training sample, the number of non-ignored targets in the last len_rec
        positions for the A-domain losses (gt_share_a, gt_a) and the B-domain losses (gt_share_b, gt_b).
        Lets the trainer size the loss GEMMs without asking the device (no host sync)."""
        key = ("valid", len_rec, na, nb)
        cache = self.__dict__.setdefault("_cache", {})
        if key not in cache:
            f = [self.fields[i][:, -len_rec:].cpu() for i in (6, 7, 8, 9)]
            a = (f[0] != na).sum(1) + (f[2] != na).sum(1)
            b = (f[1] != nb).sum(1) + (f[3] != nb).sum(1)
            cache[key] = (torch.stack((a, b), 1), na, nb)
        return cache[key]

    def check_bounds(self, n_item: int):
        """The kernels index E[seq] and P[pos] without bounds checks; the reference would raise IndexError from
        nn.Embedding (a sequence of len_max + 1 items gives pos == len_max).  Checked once per split, on the host.
        Also reports how many sequences start with a real item: their leading query rows have no allowed key under
        the reference's inverted key-padding mask -- the reference produces NaN there with an even
        head count in eval mode, this implementation defines the row as 0."""
        seqs, poss = self.fields[0:3], self.fields[3:6]
        hi_seq = max(int(x.max()) for x in seqs) if self.length else 0
        lo_seq = min(int(x.min()) for x in seqs) if self.length else 0
        hi_pos = max(int(x.max()) for x in poss) if self.length else 0
        if lo_seq < 0 or hi_seq >= n_item:
            raise IndexError(f"{self.mode} split: item id out of range [0, {n_item}) (min {lo_seq}, max {hi_seq})")
        if hi_pos >= self.len_max:
            raise IndexError(f"{self.mode} split: position {hi_pos} >= len_max = {self.len_max} "
                             "(a sequence longer than len_max items; the reference raises IndexError here too)")
        self.n_fully_masked = int((self.fields[0][:, 0] != n_item - 1).sum()) if self.length else 0
        # every PAD token carries position 0 (dataloader.py:137-143 of the reference pads both with the same
        # offsets): what the evaluation's pad-key shortcut relies on
        self.pad_pos_zero = all(bool((p[s == n_item - 1] == 0).all()) for s, p in zip(seqs, poss))
        return self.n_fully_masked

    def to(self, device, pin: bool = False):
        """Make the whole split device-resident (or pinned) once -- 'next' row f-1."""
        if pin and torch.device(device).type == "cpu":
            self.fields = [x.pin_memory() for x in self.fields]
        else:
            self.fields = [x.to(device) for x in self.fields]
            if self.mode == "train" and len({(tuple(x.shape), x.dtype) for x in self.fields}) == 1 \
                    and self.fields[0].is_cuda:
                # one [n_fields, n, L] block; the fields become views of it, a batch is ONE slice / gather
                self.stacked = torch.stack(self.fields, 0)
                self.fields = list(self.stacked.unbind(0))
        return self

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        return tuple(x[index] for x in self.fields)


class Batch(tuple):
    """A batch tuple that can carry host-side facts about itself (``n_valid``: rows of the A / B loss GEMMs) and,
    for a device-resident training split, the one ``[14, B, L]`` tensor its fields are views of (``packed``)."""
    n_valid = None
    packed = None
    global_rows = None      # rows the ranks process together in this step (the loss normaliser under data parallelism)
    global_batch = None     # samples of the global batch (differs from global_rows when a short batch is replicated)


class BatchLoader:
    """Iterates whole batches as tuples of ``[B, ...]`` tensors sliced from ``dataset.fields``.

    With ``shuffle=True`` it draws its permutation exactly as ``DataLoader(shuffle=True)``
    does (a base seed, then a sampler seed from the default generator, then
    ``torch.randperm`` on a private generator), so a seeded run visits the samples in the
    reference's order (trainer.py:15, dataloader.py:254-255).
    """

    def __init__(self, dataset: CDSRDataset, batch_size: int, shuffle: bool = False, rank: int = 0,
                 world_size: int = 1, len_rec: int = 0, ignore=None):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle
        self.rank, self.world_size = rank, world_size
        # len_rec > 0 and ignore = (n_item_a, n_item_b): training batches carry the number of non-ignored loss rows
        self.len_rec = len_rec if ignore is not None else 0
        self.ignore = ignore

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        if self.shuffle:
            torch.empty((), dtype=torch.int64).random_()                      # DataLoader's _base_seed draw
            seed = int(torch.empty((), dtype=torch.int64).random_().item())   # RandomSampler's seed draw
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(n, generator=g)
        else:
            order = None
        dev = self.dataset.fields[0].device
        counts = None
        if self.len_rec and self.dataset.mode == "train":
            na, nb = self.ignore
            counts, na, nb = self.dataset.valid_counts(self.len_rec, na, nb)
        for lo in range(0, n, self.batch_size):
            hi = min(lo + self.batch_size, n)
            g_batch = g_rows = hi - lo
            if self.world_size > 1:                     # data-parallel: contiguous slice of the global batch
                per = (hi - lo + self.world_size - 1) // self.world_size
                if per * (self.world_size - 1) >= hi - lo:
                    # a remainder so short that some rank would get nothing: every rank takes all of it.  The
                    # all-reduced normalisers (valid-target counts, rows) come out world x larger, so each rank
                    # contributes loss / world and gradient / world and the sums over ranks are the single-
                    # process values; no rank ever runs an empty step or skips a collective
                    g_rows = (hi - lo) * self.world_size
                else:
                    lo, hi = min(lo + self.rank * per, hi), min(lo + (self.rank + 1) * per, hi)
            stacked = getattr(self.dataset, "stacked", None)
            if order is None:
                if stacked is not None:
                    packed = stacked[:, lo:hi]
                    out = Batch(packed.unbind(0))
                    out.packed = packed
                else:
                    out = Batch(x[lo:hi] for x in self.dataset.fields)
                if counts is not None:
                    out.n_valid = (tuple(int(v) for v in counts[lo:hi].sum(0)), na, nb)
            else:
                idx = order[lo:hi].to(dev)
                if stacked is not None:
                    packed = stacked.index_select(1, idx)
                    out = Batch(packed.unbind(0))
                    out.packed = packed
                else:
                    out = Batch(x.index_select(0, idx) for x in self.dataset.fields)
                if counts is not None:
                    out.n_valid = (tuple(int(v) for v in counts.index_select(0, order[lo:hi]).sum(0)), na, nb)
            out.global_rows, out.global_batch = g_rows, g_batch
            yield out


def count_item(path: str) -> int:
    with open(path, "r", encoding="utf-8") as f:
        return sum(1 for _ in f)


def get_dataloader(args):
    """Reference signature (dataloader.py:245-272): sets args.n_item_a/b, n_item, idx_pad and
    returns (train, val, test) loaders."""
    p = args.path_raw if getattr(args, "use_raw", False) else args.path_data
    args.n_item_a = count_item(join(p, "items_a.txt"))
    args.n_item_b = count_item(join(p, "items_b.txt"))
    args.n_item = args.n_item_a + args.n_item_b + 1
    args.idx_pad = args.n_item - 1
    rank, world = getattr(args, "rank", 0), getattr(args, "world_size", 1)
    train = BatchLoader(CDSRDataset(args, "train"), args.batch_size, shuffle=True, rank=rank, world_size=world,
                        len_rec=getattr(args, "len_rec", 0), ignore=(args.n_item_a, args.n_item_b))
    val = BatchLoader(CDSRDataset(args, "val"), args.batch_size_eval)
    test = BatchLoader(CDSRDataset(args, "test"), args.batch_size_eval)
    return train, val, test
