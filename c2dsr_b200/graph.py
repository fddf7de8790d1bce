"""Item-item adjacencies of the cross-domain graph (reference: utils/graph.py).

``preprocess_graph`` / ``make_graph`` keep the reference's signatures and return torch
sparse COO fp32 ``[N, N]`` tensors, so pickled ``graph.pkl`` files interchange.  The
model does not multiply with the COO form: ``CsrGraph`` holds the int32 CSR of A (forward
SpMM) and of A^T (backward SpMM) that the CUDA kernels consume.
"""
from __future__ import annotations

import pickle
from os.path import join
from typing import Sequence, Tuple

import numpy as np
import torch

from .dataloader import read_raw


def transition_edges(seqs: Sequence[Sequence[int]], n_item_a: int) -> Tuple[np.ndarray, np.ndarray]:
    """Directed prev->next transitions, vectorised.

    Returns (shared [E,2], specific [E',2]).  Shared: consecutive items of the mixed
    sequence; specific: consecutive items of the same domain (utils/graph.py:55-81).
    Duplicates are kept -- the reference's de-dup sets are never filled.
    """
    lens = np.fromiter((len(s) for s in seqs), np.int64, len(seqs))
    if lens.sum() == 0:
        z = np.zeros((0, 2), np.int64)
        return z, z.copy()
    items = np.concatenate([np.asarray(s, np.int64) for s in seqs if len(s)])
    sid = np.repeat(np.arange(len(seqs)), lens)

    def consecutive(it, sd):
        same = sd[1:] == sd[:-1]
        return np.stack((it[:-1][same], it[1:][same]), 1)

    dom_a = items < n_item_a
    shared = consecutive(items, sid)
    specific = np.concatenate((consecutive(items[dom_a], sid[dom_a]), consecutive(items[~dom_a], sid[~dom_a])))
    return shared, specific


def normalised_coo(edges: np.ndarray, n: int):
    """Sum duplicate edges, then divide each row by its sum: D^-1 A (utils/graph.py:10-17,84-90).
    Returns (row, col, val) sorted by (row, col)."""
    if len(edges) == 0:
        z = np.zeros(0, np.int64)
        return z, z.copy(), np.zeros(0, np.float32)
    key, cnt = np.unique(edges[:, 0] * n + edges[:, 1], return_counts=True)
    row, col = key // n, key % n
    cnt = cnt.astype(np.float32)
    inv = np.zeros(n, np.float32)
    rowsum = np.bincount(row, weights=cnt, minlength=n).astype(np.float32)
    nz = rowsum > 0
    inv[nz] = np.float32(1.0) / rowsum[nz]
    return row, col, (inv[row] * cnt).astype(np.float32)


def _to_sparse(coo, n: int) -> torch.Tensor:
    row, col, val = coo
    idx = torch.from_numpy(np.stack((row, col)).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(val), (n, n))


def preprocess_graph(args, filename: str):
    """Reference signature utils/graph.py:33 -> (adj_share, adj_specific) sparse COO on CPU."""
    seqs = read_raw(filename)
    shared, specific = transition_edges(seqs, args.n_item_a)
    return (_to_sparse(normalised_coo(shared, args.n_item), args.n_item),
            _to_sparse(normalised_coo(specific, args.n_item), args.n_item))


def make_graph_device(args, filename: str):
    """--device_graph: transitions from the raw log (vectorised on the host), then summed duplicates, row
    normalisation, CSR and CSR^T on the GPU (CsrGraph.from_edges).  Returns two CsrGraph objects, which
    C2DSR accepts in place of the reference's COO tensors.  Same arrays, bit for bit, as make_graph()."""
    shared, specific = transition_edges(read_raw(filename), args.n_item_a)
    return (CsrGraph.from_edges(shared, args.n_item, args.device),
            CsrGraph.from_edges(specific, args.n_item, args.device))


def make_graph(args, filename: str):
    """Reference signature utils/graph.py:99-109."""
    if getattr(args, "use_raw", False):
        adj_share, adj_specific = preprocess_graph(args, filename)
        if getattr(args, "save_processed", True):
            with open(join(args.path_data, "graph.pkl"), "wb") as f:
                pickle.dump((adj_share, adj_specific), f)
    else:
        with open(join(args.path_data, "graph.pkl"), "rb") as f:
            adj_share, adj_specific = pickle.load(f)
    return adj_share.to(args.device), adj_specific.to(args.device)


class CsrGraph:
    """int32 CSR of A and of A^T for one adjacency, resident on ``device``.

    Built once from the reference's COO tensor (entries may be unsorted / duplicated:
    duplicates are summed, as ``torch.spmm`` would).  Row entries are column-sorted so the
    per-row summation order is fixed.
    """

    def __init__(self, adj: torch.Tensor, device=None):
        adj = adj.coalesce() if adj.is_sparse else adj.to_sparse().coalesce()
        n = adj.shape[0]
        idx, val = adj.indices().cpu(), adj.values().cpu().float()
        self.n, self.nnz = n, int(val.numel())
        device = device if device is not None else adj.device
        self.rowptr, self.col, self.val = self._csr(idx[0], idx[1], val, n, device)
        self.t_rowptr, self.t_col, self.t_val = self._csr(idx[1], idx[0], val, n, device)
        self.long_rows = self._long(self.rowptr)
        self.t_long_rows = self._long(self.t_rowptr)

    LONG_ROW = 256          # = c2dsr_spmm_long_row_threshold(): rows above it get a whole CTA in the SpMM

    @classmethod
    def from_edges(cls, edges, n: int, device):
        """Build A = D^-1 (summed transitions) and A^T on the device straight from raw (src, dst) pairs
        (c2dsr_graph_build: radix sort, run-length unique, bit-exact normalisation) -- the device-side
        replacement of utils/graph.py:33-96 for large logs; same arrays as CsrGraph(adj) gives for the
        reference's COO tensor."""
        from ._cabi import call, ptr, query, stream
        device = torch.device(device)
        e = torch.as_tensor(edges)
        src = e[:, 0].to(device=device, dtype=torch.int32).contiguous()
        dst = e[:, 1].to(device=device, dtype=torch.int32).contiguous()
        m = int(src.numel())
        self = cls.__new__(cls)
        self.n = n
        ws = torch.empty(query("c2dsr_graph_build_workspace_bytes", m, n), dtype=torch.uint8, device=device)
        out = []
        for transpose in (0, 1):
            rowptr = torch.empty(n + 1, dtype=torch.int32, device=device)
            col = torch.empty(max(m, 1), dtype=torch.int32, device=device)
            val = torch.empty(max(m, 1), dtype=torch.float32, device=device)
            nnz = torch.zeros(1, dtype=torch.int32, device=device)
            call("c2dsr_graph_build", ptr(src), ptr(dst), m, n, transpose, ptr(rowptr), ptr(col), ptr(val), ptr(nnz),
                 ptr(ws), ws.numel(), stream())
            k = int(nnz.item())                       # offline: one host sync per orientation
            out.append((rowptr, col[:k].clone(), val[:k].clone()))
        (self.rowptr, self.col, self.val), (self.t_rowptr, self.t_col, self.t_val) = out
        self.nnz = int(self.col.numel())
        self.long_rows = self._long(self.rowptr.cpu().long())
        self.t_long_rows = self._long(self.t_rowptr.cpu().long())
        for name in ("long_rows", "t_long_rows"):
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.to(device))
        return self

    @classmethod
    def _long(cls, rowptr):
        deg = (rowptr[1:] - rowptr[:-1])
        rows = torch.nonzero(deg > cls.LONG_ROW).view(-1).to(torch.int32)
        return rows.contiguous() if rows.numel() else None

    @property
    def fwd(self):
        return self.rowptr, self.col, self.val, self.long_rows

    @property
    def bwd(self):
        return self.t_rowptr, self.t_col, self.t_val, self.t_long_rows

    def to(self, device):
        for name in ("rowptr", "col", "val", "t_rowptr", "t_col", "t_val", "long_rows", "t_long_rows"):
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.to(device))
        return self

    @staticmethod
    def _csr(row, col, val, n, device):
        order = torch.argsort(row * n + col, stable=True)
        row, col, val = row[order], col[order], val[order]
        rowptr = torch.zeros(n + 1, dtype=torch.int64)
        rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
        return (rowptr.to(torch.int32).to(device), col.to(torch.int32).to(device).contiguous(),
                val.to(device).contiguous())
