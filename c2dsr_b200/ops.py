"""Host-side operators of the hot path: thin ``torch.autograd.Function`` wrappers that hand raw
device pointers to the C-ABI kernels (include/c2dsr_b200.h).  No arithmetic happens here except
allocation of outputs; all functions raise if the CUDA library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import torch

from . import _cabi
from ._cabi import LayerWeights, call, ptr, query, seed_tag, stream, workspace

F32, I64, I32 = torch.float32, torch.int64, torch.int32
LN_EPS = 1e-8          # models/encoders.py:24-27 of the reference


# parameter storage address -> tensor its gradient should be written into (data-parallel training over peer
# memory: the gradient block every rank can read, dist.ShardedStep._setup_peer).  Empty otherwise.
GRAD_SINKS = {}
# parameter storage address -> callable run (on the producing stream) as soon as that sink holds the complete
# gradient of the step: lets the optimiser pipeline of an embedding table start inside the backward
GRAD_READY = {}


# Sinks and ready-callbacks assume that the parameter receives exactly ONE gradient contribution in the backward:
# true for the Trainer's own step (forward_all: one branch per table; one loss node per classifier), which arms them
# around its ``loss.backward()``.  Any other backward (model API used directly, tests) never sees them.
ARMED = [False]


def grad_sink(param_ptr):
    """A fresh alias of the registered sink (autograd may then adopt it as ``.grad`` without a copy), or None."""
    t = GRAD_SINKS.get(param_ptr) if ARMED[0] else None
    return None if t is None else t.view_as(t)


def forget_grad_hooks(ptrs):
    for q in ptrs:
        GRAD_SINKS.pop(q, None)
        GRAD_READY.pop(q, None)


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous() if t.dtype == F32 else t.float().contiguous()


# ------------------------------------------------------------------------------------------------
# K2: GCN propagation (models/encoders.py:42-48)
# ------------------------------------------------------------------------------------------------
def mark_rows(ids, n_rows: int):
    """uint8 [n_rows] with 1 at every id that occurs in ``ids`` (int64, any shape)."""
    ids = ids.contiguous()
    mask = torch.empty(n_rows, dtype=torch.uint8, device=ids.device)
    call("c2dsr_mark_rows", ptr(ids, I64), ids.numel(), n_rows, ptr(mask), stream())
    return mask


def spmm(csr, X, Y=None, Z=None, out=None, alpha=1.0, beta=0.0, gamma=0.0, drop_mode=0, p=0.0, seed=0, tag=0,
         out_need=None, x_nz=None):
    """csr = (rowptr, col, val, long_rows or None) as held by graph.CsrGraph (.fwd / .bwd).  ``out_need`` /
    ``x_nz``: optional uint8 row masks (see the header): rows not wanted are left zero, all-zero rows of X are
    skipped."""
    rowptr, col, val, long_rows = csr
    n, d = X.shape
    out = torch.empty_like(X) if out is None else out
    call("c2dsr_spmm", ptr(rowptr, I32), ptr(col, I32), ptr(val, F32), ptr(long_rows),
         0 if long_rows is None else long_rows.numel(), ptr(X, F32), ptr(Y), ptr(Z), ptr(out, F32), n, d, alpha, beta,
         gamma, drop_mode, p, seed, seed_tag(seed, tag), ptr(out_need), ptr(x_nz), stream())
    return out


class GCNFn(torch.autograd.Function):
    """hi = mean([E, A drop(E), A drop(A drop(E)), ...]) with A in CSR; backward uses the CSR of A^T."""

    @staticmethod
    def forward(ctx, E, graph, n_gnn: int, p: float, seed: int, tag: int):
        E = _f(E)
        ctx.graph, ctx.n_gnn, ctx.p, ctx.seed, ctx.tag = graph, n_gnn, p, seed, tag
        return gcn_forward(E, graph, n_gnn, p, seed, tag)

    @staticmethod
    def backward(ctx, d_hi):
        g, k, p, seed, tag = ctx.graph, ctx.n_gnn, ctx.p, ctx.seed, ctx.tag
        d_hi = _f(d_hi)
        if k == 0:
            return d_hi, None, None, None, None, None
        return gcn_backward(g, d_hi, k, p, seed, tag), None, None, None, None, None


def gcn_forward(E, g, n_gnn: int, p: float, seed: int, tag: int, need=None):
    """hi = mean([E, A drop(E), A drop(A drop(E)), ...]) (models/encoders.py:42-48), n_gnn SpMMs.  ``need``
    (uint8 row mask): only those rows of hi are wanted -- the last product skips the others (left zero)."""
    if n_gnn == 0:
        return E.clone()
    c = 1.0 / (n_gnn + 1)
    h, acc = E, E
    for j in range(1, n_gnn + 1):
        t = tag * 16 + j
        if j == n_gnn:
            return spmm(g.fwd, h, Y=acc, alpha=c, beta=c, drop_mode=1, p=p, seed=seed, tag=t, out_need=need)
        h = spmm(g.fwd, h, drop_mode=1, p=p, seed=seed, tag=t)
        nxt = torch.empty_like(h)
        call("c2dsr_axpby", ptr(h), ptr(acc), ptr(nxt), h.numel(), 1.0, 1.0, stream())
        acc = nxt


def gcn_backward(g, d_hi, k: int, p: float, seed: int, tag: int, direct: bool = False, pad_idx: int = -1,
                 nz=None, out=None):
    """Gradient w.r.t. E of hi = GCN(E) (k >= 1 hops) given d_hi:  g_k = c d_hi,  g_{j-1} = c d_hi + m_j .* (A^T g_j).
    ``direct=True`` folds in the gradient of the branch's direct look-up E[seq] as well, which equals d_hi on every
    row except the pad row (``nn.Embedding(padding_idx)`` blocks it there): the last product then uses
    beta = c + 1 and the pad row is corrected, so neither a second dense [N, d] gradient nor the add of the two
    is ever materialised.  ``nz`` (uint8 row mask): rows of d_hi outside it are exactly zero (items the batch
    did not touch) and the first product skips them.  ``out``: where the last product writes the gradient."""
    c = 1.0 / (k + 1)
    extra = 1.0 if direct else 0.0
    cur = spmm(g.bwd, d_hi, Y=d_hi, alpha=c, beta=c + (extra if k == 1 else 0.0), drop_mode=2, p=p, seed=seed,
               tag=tag * 16 + k, x_nz=nz, out=out if k == 1 else None)
    for j in range(k - 1, 0, -1):
        cur = spmm(g.bwd, cur, Y=d_hi, alpha=1.0, beta=c + (extra if j == 1 else 0.0), drop_mode=2, p=p, seed=seed,
                   tag=tag * 16 + j, out=out if j == 1 else None)
    if direct:
        cur[pad_idx].sub_(d_hi[pad_idx])
    return cur


# ------------------------------------------------------------------------------------------------
# K1: branch input (models/C2DSR.py:65-71, models/encoders.py:30-31)
# ------------------------------------------------------------------------------------------------
def _gather_forward(hi, E, P, seq, pos, scale, p, seed, tag):
    d = E.shape[1]
    x = torch.empty(*seq.shape, d, device=E.device, dtype=F32)
    call("c2dsr_gather_fwd", ptr(hi), ptr(E), ptr(P), ptr(seq, I64), ptr(pos, I64), ptr(x), seq.numel(),
         d, scale, p, seed, seed_tag(seed, tag), stream())
    return x


def _gather_backward(dx, seq, pos, n_rows, d, p_shape, scale, pad_idx, p, seed, tag, want_dE=True):
    d_hi = torch.zeros(n_rows, d, device=dx.device, dtype=F32)
    d_E = torch.zeros(n_rows, d, device=dx.device, dtype=F32) if want_dE else None
    d_P = torch.zeros(p_shape, device=dx.device, dtype=F32)
    n = seq.numel()
    nb = query("c2dsr_gather_bwd_workspace_bytes", n, d, n_rows, p_shape[0])
    ws = workspace.get(nb, dx.device)
    call("c2dsr_gather_bwd", ptr(dx), ptr(seq), ptr(pos), ptr(d_hi), ptr(d_E), ptr(d_P), n, d, n_rows, p_shape[0],
         pad_idx, scale, p, seed, seed_tag(seed, tag), ptr(ws), ws.numel(), stream())
    return d_hi, d_E, d_P


class GatherFn(torch.autograd.Function):
    """x = drop(sqrt(d) * (hi[seq] + E[seq]) + P[pos]);  seq, pos int64 of any shape -> [..., d]."""

    @staticmethod
    def forward(ctx, hi, E, P, seq, pos, scale: float, pad_idx: int, p: float, seed: int, tag: int):
        hi, E, P = _f(hi), _f(E), _f(P)
        seq_c, pos_c = seq.contiguous(), pos.contiguous()
        x = _gather_forward(hi, E, P, seq_c, pos_c, scale, p, seed, tag)
        ctx.save_for_backward(seq_c, pos_c)
        ctx.cfg = (hi.shape, P.shape, scale, pad_idx, p, seed, tag)
        return x

    @staticmethod
    def backward(ctx, dx):
        seq, pos = ctx.saved_tensors
        (n_rows, d), p_shape, scale, pad_idx, p, seed, tag = ctx.cfg
        d_hi, d_E, d_P = _gather_backward(_f(dx), seq, pos, n_rows, d, p_shape, scale, pad_idx, p, seed, tag)
        return d_hi, d_E, d_P, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# K3: encoder (models/encoders.py:23-33)
# ------------------------------------------------------------------------------------------------
_LAYER_FIELDS = ("in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b",
                 "ln1_w", "ln1_b", "ln2_w", "ln2_b")


def _layer_table(tensors: Sequence[torch.Tensor], n_layers: int):
    arr = (LayerWeights * max(n_layers, 1))()
    for l in range(n_layers):
        for k, name in enumerate(_LAYER_FIELDS):
            setattr(arr[l], name, ptr(tensors[12 * l + k], F32))
    return arr


def _encoder_forward(x, seq, w, n_head, pad_idx, norm_first, p, seed, tag, dense_passes):
    n_seq, L, d = x.shape
    n_layers = (len(w) - 2) // 12
    T = n_seq * L
    saved = torch.empty(query("c2dsr_encoder_saved_floats", T, d, n_head, n_layers), device=x.device, dtype=F32)
    out = torch.empty_like(x)
    ws = workspace.get(query("c2dsr_encoder_workspace_bytes", T, d, n_head, dense_passes), x.device)
    table = _layer_table(w, n_layers)
    call("c2dsr_encoder_fwd", C.addressof(table), n_layers, ptr(w[-2]), ptr(w[-1]), ptr(x), ptr(seq, I64), n_seq,
         L, d, n_head, pad_idx, int(norm_first), dense_passes, LN_EPS, p, seed, seed_tag(seed, tag), ptr(out),
         ptr(saved), ptr(ws), ws.numel(), stream())
    return out, saved


def _encoder_backward(d_out, saved, seq, w, cfg):
    n_seq, L, d, n_head, pad_idx, norm_first, p, seed, tag, n_layers, dense_passes = cfg
    # one zero-filled buffer, one view per weight (a single memset instead of 12 * n_layers + 2)
    sizes = [(t.numel() + 3) // 4 * 4 for t in w]
    flat = torch.zeros(sum(sizes), device=d_out.device, dtype=F32)
    grads, off = [], 0
    for t, n in zip(w, sizes):
        grads.append(flat[off:off + t.numel()].view(t.shape))
        off += n
    dx = torch.empty(n_seq, L, d, device=d_out.device, dtype=F32)
    ws = workspace.get(query("c2dsr_encoder_workspace_bytes", n_seq * L, d, n_head, dense_passes), d_out.device)
    wt, gt = _layer_table(w, n_layers), _layer_table(grads, n_layers)
    call("c2dsr_encoder_bwd", C.addressof(wt), C.addressof(gt), n_layers, ptr(w[-2]), ptr(grads[-2]),
         ptr(grads[-1]), ptr(d_out), ptr(seq), n_seq, L, d, n_head, pad_idx, int(norm_first), dense_passes, LN_EPS,
         p, seed, seed_tag(seed, tag), ptr(saved), ptr(dx), ptr(ws), ws.numel(), stream())
    return dx, grads


class EncoderFn(torch.autograd.Function):
    """x [n_seq, L, d], seq [n_seq, L] -> encoder output.  ``weights`` = 12 tensors per layer in the
    order of _LAYER_FIELDS, then the final LayerNorm weight and bias."""

    @staticmethod
    def forward(ctx, x, seq, n_head: int, pad_idx: int, norm_first: bool, p: float, seed: int, tag: int,
                dense_passes: int, *weights):
        x = _f(x)
        seq = seq.contiguous()
        w = [_f(t) for t in weights]
        out, saved = _encoder_forward(x, seq, w, n_head, pad_idx, norm_first, p, seed, tag, dense_passes)
        ctx.save_for_backward(saved, seq, *w)
        ctx.cfg = (*x.shape, n_head, pad_idx, norm_first, p, seed, tag, (len(w) - 2) // 12, dense_passes)
        return out

    @staticmethod
    def backward(ctx, d_out):
        saved, seq, *w = ctx.saved_tensors
        dx, grads = _encoder_backward(_f(d_out), saved, seq, w, ctx.cfg)
        return (dx, None, None, None, None, None, None, None, None, *grads)


@torch.no_grad()
def branch_select(hi, E, P, seq, pos, sel, scale: float, pad_idx: int, n_head: int, norm_first: bool, dense_passes: int,
                  weights):
    """Evaluation form of one branch for a one-layer encoder: h[b, sel[b], :] only (no autograd, no dropout).
    Gather for every token (keys and values need them all), then c2dsr_encoder_fwd_select."""
    hi, E, P = _f(hi), _f(E), _f(P)
    seq, pos, sel = seq.contiguous(), pos.contiguous(), sel.contiguous()
    w = [_f(t) for t in weights]
    n_seq, L = seq.shape
    d = E.shape[1]
    x = _gather_forward(hi, E, P, seq, pos, scale, 0.0, 0, 0)
    out = torch.empty(n_seq, d, device=E.device, dtype=F32)
    ws = workspace.get(query("c2dsr_encoder_select_workspace_bytes", n_seq, L, d, dense_passes), E.device)
    table = _layer_table(w, 1)
    call("c2dsr_encoder_fwd_select", C.addressof(table), 1, ptr(w[-2]), ptr(w[-1]), ptr(x), ptr(seq, I64), ptr(sel, I64),
         n_seq, L, d, n_head, pad_idx, int(norm_first), dense_passes, LN_EPS, ptr(out), ptr(ws), ws.numel(), stream())
    return out


@torch.no_grad()
def branch_padkeys(hi, E, P, seq, pos, sel, scale: float, pad_idx: int, n_head: int, norm_first: bool,
                   dense_passes: int, weights, cache=None):
    """Evaluation form of one branch when every PAD token has position 0 (the preprocessor's invariant, checked by
    CDSRDataset.check_bounds): h[b, sel[b], :] without projecting or attending over the B * L tokens at all -- the
    reference's inverted key-padding mask lets a query see PAD keys only, and those are all the same
    row.  One gather of B + 1 tokens (the selected ones and one PAD token), then c2dsr_encoder_fwd_padkeys.  ``cache``
    (a dict owned by the caller, emptied whenever the propagated tables or the weights change) keeps the PAD token's
    attention-block output, which does not depend on the batch."""
    hi, E, P = _f(hi), _f(E), _f(P)
    seq, pos, sel = seq.contiguous(), pos.contiguous(), sel.contiguous()
    w = [_f(t) for t in weights]
    n_seq, L = seq.shape
    d = E.shape[1]
    x = torch.empty(n_seq + 1, d, device=E.device, dtype=F32)          # the last row is the PAD token
    call("c2dsr_gather_select_fwd", ptr(hi), ptr(E), ptr(P), ptr(seq, I64), ptr(pos, I64), ptr(sel, I64), ptr(x), n_seq, L,
         d, scale, pad_idx, stream())
    table = _layer_table(w, 1)
    y_pad = cache.get("y_pad") if cache is not None else None
    if y_pad is None:
        y_pad = torch.empty(d, device=E.device, dtype=F32)
        ws = workspace.get(32 * d + (16 << 20) + 1024, E.device)
        call("c2dsr_encoder_padkeys_prepare", C.addressof(table), 1, ptr(x[n_seq:]), d, int(norm_first), LN_EPS, ptr(y_pad),
             ptr(ws), ws.numel(), stream())
        if cache is not None and not torch.cuda.is_current_stream_capturing():
            cache["y_pad"] = y_pad
    out = torch.empty(n_seq, d, device=E.device, dtype=F32)
    ws = workspace.get(query("c2dsr_encoder_padkeys_workspace_bytes", n_seq, d, dense_passes), E.device)
    call("c2dsr_encoder_fwd_padkeys", C.addressof(table), 1, ptr(w[-2]), ptr(w[-1]), ptr(x), ptr(y_pad), ptr(seq, I64),
         ptr(sel, I64), n_seq, L, d, n_head, pad_idx, int(norm_first), dense_passes, LN_EPS, ptr(out), ptr(ws),
         ws.numel(), stream())
    return out


class BranchSetFn(torch.autograd.Function):
    """Several independent branches (K1 gather + K3 encoder each) as ONE autograd node that forks them
    onto CUDA streams and joins before returning, in forward and in backward.  One branch alone is too
    small for 148 SMs (60-180 GEMM tiles); the autograd engine only ever sees the caller's stream, so
    no cross-stream bookkeeping is left to it.

    ``specs``: per branch a dict(seq, pos, scale, pad, p, seed, gather_tag, encoder_tag, n_head,
    norm_first, dense_passes, n_w, gcn); ``streams``: per branch a torch.cuda.Stream or None (= the caller's
    stream); ``flat``: per branch hi, E, P and the n_w encoder weights.  ``gcn`` = (graph, n_gnn, p, seed,
    tag) says that ``hi`` (passed detached) is GCN(E) of the same step: the backward then continues through
    the propagation on the branch's own stream and returns the whole gradient of E (see gcn_backward).  With
    ``gcn`` given and ``hi`` = None the propagation itself is also computed here, on the branch's stream, and
    returned as an extra (non-differentiable) output after the branch outputs."""

    @staticmethod
    def forward(ctx, specs, streams, *flat):
        cur = torch.cuda.current_stream()
        outs, keep, off = [None] * len(specs), [], 0
        parts = []
        for sp in specs:
            n = 3 + sp["n_w"]
            parts.append([None if t is None else _f(t) for t in flat[off:off + n]])
            off += n
        order = sorted(range(len(specs)), key=lambda i: streams[i] is None)      # side streams first
        his = [None] * len(specs)
        for i in order:
            sp, (hi, E, P, *w) = specs[i], parts[i]
            seq, pos = sp["seq"].contiguous(), sp["pos"].contiguous()
            st = streams[i] or cur
            if st is not cur:
                st.wait_stream(cur)
            with torch.cuda.stream(st):
                if hi is None:                       # propagate on this branch's stream, only the rows the batch reads
                    graph, k, gp, gseed, gtag = sp["gcn"]
                    hi = his[i] = gcn_forward(E, graph, k, gp, gseed, gtag, need=mark_rows(seq, E.shape[0]))
                x = _gather_forward(hi, E, P, seq, pos, sp["scale"], sp["p"], sp["seed"], sp["gather_tag"])
                out, saved = _encoder_forward(x, seq, w, sp["n_head"], sp["pad"], sp["norm_first"], sp["p"],
                                              sp["seed"], sp["encoder_tag"], sp["dense_passes"])
                del x
            outs[i] = out
            keep.append((i, saved, seq, pos, hi.shape, P.shape, tuple(out.shape), len(w)))
        for st in streams:
            if st is not None:
                cur.wait_stream(st)
        ctx.specs, ctx.streams = specs, streams
        ctx.meta = [(i, seq, pos, hs, ps, os, nw) for i, _, seq, pos, hs, ps, os, nw in keep]
        ctx.save_for_backward(*[k[1] for k in keep], *[t for prt in parts for t in prt[3:]])
        extra = tuple(h for h in his if h is not None)
        ctx.mark_non_differentiable(*extra)
        return tuple(outs) + extra

    @staticmethod
    def backward(ctx, *d_outs):
        specs, streams = ctx.specs, ctx.streams
        nb = len(specs)
        saved_all = ctx.saved_tensors
        wts, off = [], nb
        for sp in specs:
            wts.append(list(saved_all[off:off + sp["n_w"]]))
            off += sp["n_w"]
        cur = torch.cuda.current_stream()
        result = [None] * nb
        for slot, (i, seq, pos, hi_shape, p_shape, o_shape, nw) in enumerate(ctx.meta):
            sp, st = specs[i], streams[i] or cur
            d_out = d_outs[i]
            if d_out is None:
                d_out = torch.zeros(o_shape, device=seq.device, dtype=F32)
            d_out = _f(d_out)
            if st is not cur:
                st.wait_stream(cur)
            with torch.cuda.stream(st):
                n_seq, L, d = o_shape
                cfg = (n_seq, L, d, sp["n_head"], sp["pad"], sp["norm_first"], sp["p"], sp["seed"],
                       sp["encoder_tag"], (nw - 2) // 12, sp["dense_passes"])
                dx, grads = _encoder_backward(d_out, saved_all[slot], seq, wts[i], cfg)
                gcn = sp.get("gcn")
                d_hi, d_E, d_P = _gather_backward(dx, seq, pos, hi_shape[0], d, p_shape, sp["scale"], sp["pad"],
                                                  sp["p"], sp["seed"], sp["gather_tag"], want_dE=gcn is None)
                del dx
                if gcn is not None:
                    graph, k, gp, gseed, gtag = gcn
                    # (sink / early step only for a table that a single branch of this step reads)
                    tp = sp.get("table_ptr")
                    single = tp is not None and sum(1 for s in specs if s.get("table_ptr") == tp) == 1
                    d_E = gcn_backward(graph, d_hi, k, gp, gseed, gtag, direct=True, pad_idx=sp["pad"],
                                       nz=mark_rows(seq, hi_shape[0]), out=grad_sink(tp) if single else None)
                    ready = GRAD_READY.get(tp) if (ARMED[0] and single) else None
                    if ready is not None:
                        ready()
                    d_hi = None
            result[i] = [d_hi, d_E, d_P, *grads]
        for st in streams:
            if st is not None:
                cur.wait_stream(st)
        return (None, None, *[t for r in result for t in r])


# ------------------------------------------------------------------------------------------------
# K5: infomax (trainer.py:85-119)
# ------------------------------------------------------------------------------------------------
class InfomaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_share, hx, hy, h_neg_a, h_neg_b, W_a, W_b, bias_a, bias_b, gt_mask_a, gt_mask_b,
                inv_batch: float):
        hs = [_f(t) for t in (h_share, hx, hy, h_neg_a, h_neg_b)]
        B, L, d = hs[0].shape
        Wa, Wb = _f(W_a).view(d, d), _f(W_b).view(d, d)
        ma, mb = gt_mask_a.contiguous(), gt_mask_b.contiguous()
        dev = hs[0].device
        pooled = torch.empty(6, B, d, device=dev, dtype=F32)
        U = torch.empty(4, B, d, device=dev, dtype=F32)
        sims = torch.empty(4, B, device=dev, dtype=F32)
        loss = torch.empty((), device=dev, dtype=F32)
        ws = workspace.get(query("c2dsr_infomax_workspace_bytes", B, d), dev)
        call("c2dsr_infomax_fwd", *(ptr(t) for t in hs), ptr(ma, I64), ptr(mb, I64), ptr(Wa), ptr(Wb),
             ptr(bias_a), ptr(bias_b), B, L, d, inv_batch, ptr(pooled), ptr(U), ptr(sims), ptr(loss), ptr(ws),
             ws.numel(), stream())
        ctx.save_for_backward(pooled, U, sims, ma, mb, Wa, Wb)
        ctx.cfg = (B, L, d, inv_batch, bias_a is not None, W_a.shape)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        pooled, U, sims, ma, mb, Wa, Wb = ctx.saved_tensors
        B, L, d, inv_batch, has_bias, w_shape = ctx.cfg
        dev = pooled.device
        d_loss = _f(d_loss).reshape(1)
        dh = [torch.empty(B, L, d, device=dev, dtype=F32) for _ in range(5)]
        dWa, dWb = torch.zeros(d, d, device=dev, dtype=F32), torch.zeros(d, d, device=dev, dtype=F32)
        dba = torch.zeros(1, device=dev, dtype=F32) if has_bias else None
        dbb = torch.zeros(1, device=dev, dtype=F32) if has_bias else None
        ws = workspace.get(query("c2dsr_infomax_workspace_bytes", B, d), dev)
        call("c2dsr_infomax_bwd", ptr(d_loss), ptr(pooled), ptr(U), ptr(sims), ptr(ma), ptr(mb), ptr(Wa), ptr(Wb), B,
             L, d, inv_batch, *(ptr(t) for t in dh), ptr(dWa), ptr(dWb), ptr(dba), ptr(dbb), ptr(ws), ws.numel(),
             stream())
        return (*dh, dWa.view(w_shape), dWb.view(w_shape), dba, dbb, None, None, None)


# ------------------------------------------------------------------------------------------------
# K4a: classifier + cross-entropy (trainer.py:131-152)
# ------------------------------------------------------------------------------------------------
_GEMM_WS = 64 << 20


def gemm(ta, tb, M, N, K, A, lda, B, ldb, Cm, ldc, alpha=1.0, beta=0.0, bias=None, act=0):
    ws = workspace.get(_GEMM_WS, Cm.device)
    call("c2dsr_gemm", ta, tb, M, N, K, alpha, ptr(A), lda, ptr(B), ldb, beta, ptr(Cm), ldc, ptr(bias), act, 0.0, 0,
         0, ptr(ws), ws.numel(), stream())
    return Cm


class ScoreCEFn(torch.autograd.Function):
    """sum_m rowscale[m] * CE([H W^T + b | Hpad wpad^T + bpad][m], gt[m]) with ignore_index = N.

    H, Hpad [M, d]; W [N, d]; gt [M] int64 in [0, N]; rowscale [M] fp32 (device)."""

    @staticmethod
    def forward(ctx, H, Hpad, W, b, wpad, bpad, gt, rowscale):
        H, Hpad, W, b, wpad, bpad = (_f(t) for t in (H, Hpad, W, b, wpad, bpad))
        gt, rowscale = gt.contiguous(), _f(rowscale)
        M, d = H.shape
        N = W.shape[0]
        dev = H.device
        ldz = query("c2dsr_score_ldz", N)
        zpad = torch.empty(M, device=dev, dtype=F32)
        gemm(0, 1, M, 1, d, Hpad, d, wpad, d, zpad, 1, bias=bpad)
        Z = torch.empty(M, ldz, device=dev, dtype=F32)
        lse = torch.empty(M, device=dev, dtype=F32)
        loss_row = torch.empty(M, device=dev, dtype=F32)
        ws = workspace.get(_GEMM_WS, dev)
        call("c2dsr_score_ce_fwd", ptr(H), ptr(W), ptr(b), ptr(zpad), ptr(gt, I64), M, N, d, ptr(Z), ptr(lse),
             ptr(loss_row), ptr(ws), ws.numel(), stream())
        loss = torch.empty((), device=dev, dtype=F32)
        call("c2dsr_wsum", ptr(loss_row), ptr(rowscale), M, ptr(loss), stream())
        ctx.save_for_backward(H, Hpad, W, wpad, gt, rowscale, zpad, Z, lse)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        H, Hpad, W, wpad, gt, rowscale, zpad, Z, lse = ctx.saved_tensors
        M, d = H.shape
        N = W.shape[0]
        dev = H.device
        coef = (rowscale * d_loss).contiguous()
        dH = torch.empty(M, d, device=dev, dtype=F32)
        dW = torch.zeros(N, d, device=dev, dtype=F32)
        db = torch.zeros(N, device=dev, dtype=F32)
        dzpad = torch.empty(M, device=dev, dtype=F32)
        ws = workspace.get(_GEMM_WS, dev)
        call("c2dsr_score_ce_bwd", ptr(H), ptr(W), ptr(zpad), ptr(gt), ptr(lse), ptr(coef), M, N, d, ptr(Z), ptr(dH),
             ptr(dW), ptr(db), ptr(dzpad), ptr(ws), ws.numel(), stream())
        dHpad = torch.empty(M, d, device=dev, dtype=F32)
        gemm(0, 0, M, d, 1, dzpad, 1, wpad, d, dHpad, d)                 # outer product dzpad (x) wpad
        dwpad = torch.empty(1, d, device=dev, dtype=F32)
        gemm(1, 0, 1, d, M, dzpad, 1, Hpad, d, dwpad, d)                 # dzpad^T Hpad
        dbpad = torch.empty(1, device=dev, dtype=F32)
        call("c2dsr_wsum", ptr(dzpad), None, M, ptr(dbpad), stream())
        return dH, dHpad, dW, db, dwpad, dbpad, None, None


def gemm_tc(ta, tb, M, N, K, A, lda, B, ldb, Cm, ldc, beta=0.0, bias=None, act=0, passes=3):
    """C = act(op(A) op(B) + bias) + beta C on the tcgen05 path (same operand conventions as gemm())."""
    ws = workspace.get(query("c2dsr_gemm_tc_workspace_bytes", M, N, K), Cm.device)
    call("c2dsr_gemm_tc", ta, tb, M, N, K, ptr(A), lda, ptr(B), ldb, beta, ptr(Cm), ldc, ptr(bias), act, 0.0, 0, 0,
         passes, ptr(ws), ws.numel(), stream())
    return Cm


class ScoreCETcFn(torch.autograd.Function):
    """Same contract as ScoreCEFn on the tcgen05 path: logits stay in tensor memory (forward keeps
    per-tile log-sum-exp partials; backward recomputes, stores dZ as bf16 hi/lo and runs two more GEMMs)."""

    @staticmethod
    def forward(ctx, H, Hpad, W, b, wpad, bpad, gt, rowscale, passes: int):
        H, Hpad, W, b, wpad, bpad = (_f(t) for t in (H, Hpad, W, b, wpad, bpad))
        gt, rowscale = gt.contiguous(), _f(rowscale)
        M, d = H.shape
        N = W.shape[0]
        dev = H.device
        zpad = torch.empty(M, device=dev, dtype=F32)
        gemm(0, 1, M, 1, d, Hpad, d, wpad, d, zpad, 1, bias=bpad)
        lse = torch.empty(M, device=dev, dtype=F32)
        loss_row = torch.empty(M, device=dev, dtype=F32)
        ws = workspace.get(query("c2dsr_score_ce_tc_workspace_bytes", M, N, d, 0), dev)
        call("c2dsr_score_ce_fwd_tc", ptr(H), ptr(W), None, None, ptr(b), ptr(zpad), ptr(gt, I64), M, N, d, passes,
             ptr(lse), ptr(loss_row), ptr(ws), ws.numel(), stream())
        loss = torch.empty((), device=dev, dtype=F32)
        call("c2dsr_wsum", ptr(loss_row), ptr(rowscale), M, ptr(loss), stream())
        ctx.save_for_backward(H, Hpad, W, b, wpad, gt, rowscale, zpad, lse)
        ctx.passes = passes
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        H, Hpad, W, b, wpad, gt, rowscale, zpad, lse = ctx.saved_tensors
        M, d = H.shape
        N = W.shape[0]
        dev = H.device
        coef = (rowscale * d_loss).contiguous()
        dH = torch.empty(M, d, device=dev, dtype=F32)
        dW = torch.zeros(N, d, device=dev, dtype=F32)
        db = torch.zeros(N, device=dev, dtype=F32)
        dzpad = torch.empty(M, device=dev, dtype=F32)
        ws = workspace.get(query("c2dsr_score_ce_tc_workspace_bytes", M, N, d, 1), dev)
        call("c2dsr_score_ce_bwd_tc", ptr(H), ptr(W), None, None, ptr(b), ptr(zpad), ptr(gt), ptr(lse), ptr(coef), M,
             N, d, ctx.passes, ptr(dH), ptr(dW), ptr(db), ptr(dzpad), ptr(ws), ws.numel(), stream())
        dHpad = torch.empty(M, d, device=dev, dtype=F32)
        gemm(0, 0, M, d, 1, dzpad, 1, wpad, d, dHpad, d)
        dwpad = torch.empty(1, d, device=dev, dtype=F32)
        gemm(1, 0, 1, d, M, dzpad, 1, Hpad, d, dwpad, d)
        dbpad = torch.empty(1, device=dev, dtype=F32)
        call("c2dsr_wsum", ptr(dzpad), None, M, ptr(dbpad), stream())
        return dH, dHpad, dW, db, dwpad, dbpad, None, None, None


class DomainLossTcFn(torch.autograd.Function):
    """The recommendation loss of one domain (trainer.py:122-152) from the encoder outputs themselves:
    h_share, h_dom [B, L, d], targets gt_share, gt_dom [B, L] (ignore = n_cls), the last R positions.  Row
    assembly (c2dsr_loss_rows_*) + the tcgen05 logits / cross-entropy (c2dsr_score_ce_*_tc) in one autograd node;
    M = number of rows sent through the GEMMs (valid rows first; 2 B R = all, like the reference)."""

    @staticmethod
    def forward(ctx, h_share, h_dom, W, b, wpad, bpad, gt_share, gt_dom, w_share, n_dom, R: int, M: int, passes: int):
        h_share, h_dom, W, b, wpad, bpad = (_f(t) for t in (h_share, h_dom, W, b, wpad, bpad))
        gt_share, gt_dom = gt_share.contiguous(), gt_dom.contiguous()
        w_share, n_dom = _f(w_share).reshape(1), _f(n_dom).reshape(1)
        B, L, d = h_share.shape
        N = W.shape[0]
        dev = h_share.device
        perm = torch.empty(2 * B * R, device=dev, dtype=I64)
        inv = torch.empty(2 * B * R, device=dev, dtype=I32)
        H = torch.empty(M, d, device=dev, dtype=F32)
        gt = torch.empty(M, device=dev, dtype=I64)
        w = torch.empty(M, device=dev, dtype=F32)
        zpad = torch.empty(M, device=dev, dtype=F32)
        call("c2dsr_loss_rows_fwd", ptr(h_share), ptr(h_dom), ptr(gt_share, I64), ptr(gt_dom, I64), B, L, R, d, N, M,
             ptr(w_share), ptr(n_dom), ptr(wpad), ptr(bpad), ptr(perm), ptr(inv), ptr(H), ptr(gt), ptr(w), ptr(zpad),
             stream())
        loss = torch.zeros((), device=dev, dtype=F32)
        lse = torch.empty(M, device=dev, dtype=F32)
        # one bf16 hi / lo split of the classifier weights per step, shared by forward and backward
        w_hi, w_lo = split_bf16(W, passes == 3) if M > 0 else (torch.empty(0, device=dev, dtype=BF16),) * 2
        if w_lo is None:
            w_lo = torch.empty(0, device=dev, dtype=BF16)
        if M > 0:
            loss_row = torch.empty(M, device=dev, dtype=F32)
            ws = workspace.get(query("c2dsr_score_ce_tc_workspace_bytes", M, N, d, 0), dev)
            call("c2dsr_score_ce_fwd_tc", ptr(H), ptr(W), ptr(w_hi), ptr(w_lo) if w_lo.numel() else None, ptr(b),
                 ptr(zpad), ptr(gt, I64), M, N, d, passes, ptr(lse), ptr(loss_row), ptr(ws), ws.numel(), stream())
            call("c2dsr_wsum", ptr(loss_row), ptr(w), M, ptr(loss), stream())
        ctx.save_for_backward(h_share, h_dom, W, b, wpad, perm, inv, H, gt, w, zpad, lse, w_hi, w_lo)
        ctx.cfg = (R, M, passes)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        h_share, h_dom, W, b, wpad, perm, inv, H, gt, w, zpad, lse, w_hi, w_lo = ctx.saved_tensors
        R, M, passes = ctx.cfg
        B, L, d = h_share.shape
        N = W.shape[0]
        dev = h_share.device
        dH = torch.empty(M, d, device=dev, dtype=F32)
        dW = grad_sink(W.data_ptr())
        dW = torch.zeros(N, d, device=dev, dtype=F32) if dW is None else dW.zero_()
        db = torch.zeros(N, device=dev, dtype=F32)
        dzpad = torch.empty(M, device=dev, dtype=F32)
        if M > 0:
            coef = (w * d_loss).contiguous()
            ws = workspace.get(query("c2dsr_score_ce_tc_workspace_bytes", M, N, d, 1), dev)
            call("c2dsr_score_ce_bwd_tc", ptr(H), ptr(W), ptr(w_hi), ptr(w_lo) if w_lo.numel() else None, ptr(b),
                 ptr(zpad), ptr(gt), ptr(lse), ptr(coef), M, N, d, passes, ptr(dH), ptr(dW), ptr(db), ptr(dzpad),
                 ptr(ws), ws.numel(), stream())
        d_share = torch.empty_like(h_share)
        d_dom = torch.empty_like(h_dom)
        dwpad = torch.empty(1, d, device=dev, dtype=F32)
        dbpad = torch.empty(1, device=dev, dtype=F32)
        ws = workspace.get(query("c2dsr_loss_rows_bwd_workspace_bytes", M, d), dev)
        call("c2dsr_loss_rows_bwd", ptr(dH), ptr(dzpad), ptr(h_share), ptr(h_dom), ptr(wpad), ptr(perm), ptr(inv), B, L,
             R, d, M, ptr(d_share), ptr(d_dom), ptr(dwpad), ptr(dbpad), ptr(ws), ws.numel(), stream())
        return d_share, d_dom, dW, db, dwpad, dbpad, None, None, None, None, None, None, None


def preprocess_train(items, offs, draws, n_item_a: int, n_item_b: int, len_max: int):
    """Device preprocessor of the training split (dataloader.py:60-161): ``items`` / ``offs`` the concatenated item
    lists, ``draws`` the host-drawn corruption ids (one per input position).  Returns (fields [n, 14, L] int64 of
    every sequence, keep [n] bool -- the reference drops the others)."""
    n = offs.numel() - 1
    fields = torch.empty(n, 14, len_max, dtype=I64, device=items.device)
    keep = torch.empty(n, dtype=torch.uint8, device=items.device)
    call("c2dsr_preprocess_train", ptr(items, I64), ptr(offs, I64), ptr(draws, I64), n, n_item_a, n_item_b, len_max,
         ptr(fields, I64), ptr(keep), stream())
    return fields, keep.bool()


def preprocess_eval(items, offs, picks, n_item_a: int, n_item_b: int, len_max: int):
    """Device preprocessor of an evaluation split (dataloader.py:163-228); ``picks`` [n, n_neg] is the host's sample
    of range(population), turned into the negative ids in place.  Returns (six [n, 6, L], four [n, 4], picks)."""
    n = offs.numel() - 1
    six = torch.empty(n, 6, len_max, dtype=I64, device=items.device)
    four = torch.empty(n, 4, dtype=I64, device=items.device)
    picks = picks.contiguous()
    call("c2dsr_preprocess_eval", ptr(items, I64), ptr(offs, I64), n, n_item_a, n_item_b, len_max, picks.shape[1],
         ptr(six, I64), ptr(four, I64), ptr(picks, I64), stream())
    return six, four, picks


def compact_rows(gt: torch.Tensor, ignore: int) -> torch.Tensor:
    """Stable partition of the row indices: rows with gt != ignore first (int64 [M])."""
    gt = gt.contiguous()
    perm = torch.empty_like(gt)
    call("c2dsr_compact_rows", ptr(gt, I64), gt.numel(), ignore, ptr(perm, I64), stream())
    return perm


def score_ce(H, Hpad, W, b, wpad, bpad, gt, rowscale, path: str = "tc", passes: int = 3):
    """Weighted cross-entropy sum over the full catalogue; path 'tc' (tcgen05) or 'ffma' (fp32 CUDA cores)."""
    if path == "tc":
        return ScoreCETcFn.apply(H, Hpad, W, b, wpad, bpad, gt, rowscale, passes)
    return ScoreCEFn.apply(H, Hpad, W, b, wpad, bpad, gt, rowscale)


# ------------------------------------------------------------------------------------------------
# K4b: scoring + rank counting (trainer.py:168-179)
# ------------------------------------------------------------------------------------------------
def score_shard(Q, W, bias) -> torch.Tensor:
    """S[n_q, lds] = Q W^T + b (fp32 FFMA path; lds = n_shard rounded up to 4)."""
    Q, W, bias = _f(Q), _f(W), _f(bias)
    n_q, d = Q.shape
    n = W.shape[0]
    lds = (n + 3) // 4 * 4
    S = torch.empty(n_q, lds, device=Q.device, dtype=F32)
    ws = workspace.get(_GEMM_WS, Q.device)
    call("c2dsr_score_shard", ptr(Q), ptr(W), ptr(bias), n_q, n, d, ptr(S), lds, ptr(ws), ws.numel(), stream())
    return S


def pick_target(S, gt, n0: int, n1: int) -> torch.Tensor:
    s_gt = torch.empty(S.shape[0], device=S.device, dtype=F32)
    call("c2dsr_pick_target", ptr(S), S.stride(0), ptr(gt.contiguous(), I64), S.shape[0], n0, n1, ptr(s_gt), stream())
    return s_gt


def rank_from_scores(S, s_gt, gt, neg: Optional[torch.Tensor], n0: int, n1: int,
                     counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """counts[i] += #{candidates j in [n0, n1), j != gt[i] : S[i, j - n0] > s_gt[i]} (int32)."""
    n_q = S.shape[0]
    if counts is None:
        counts = torch.zeros(n_q, device=S.device, dtype=I32)
    negc = None if neg is None else neg.contiguous()
    call("c2dsr_rank_from_scores", ptr(S, F32), S.stride(0), ptr(s_gt, F32), ptr(gt.contiguous(), I64),
         ptr(negc), 0 if negc is None else negc.shape[1], n_q, n0, n1, ptr(counts, I32), stream())
    return counts


# ------------------------------------------------------------------------------------------------
# K4b, tensor-core path (tcgen05): fused score GEMM + rank count
# ------------------------------------------------------------------------------------------------
BF16 = torch.bfloat16


def split_bf16(X: torch.Tensor, want_lo: bool = True, out=None):
    """x = hi + lo with hi = bf16(x), lo = bf16(x - hi): operands of the 3-pass (fp32-grade) MMA.  ``out`` =
    an earlier (hi, lo) result to overwrite in place."""
    X = _f(X)
    rows, d = X.shape
    if out is not None:
        hi, lo = out
    else:
        hi = torch.empty(rows, d, device=X.device, dtype=BF16)
        lo = torch.empty(rows, d, device=X.device, dtype=BF16) if want_lo else None
    call("c2dsr_split_bf16", ptr(X), rows, d, d, ptr(hi), ptr(lo), stream())
    return hi, lo


def score_rank_tc(Q, W_split, bias, gt, n0: int, n1: int, passes: int = 3, neg=None, s_gt=None,
                  counts=None, want_scores: bool = False, reduce_s_gt=None, n_q_limit=None):
    """Full-catalogue partial rank counts of queries Q against the catalogue shard [n0, n1).

    W_split = split_bf16(W[n0:n1]) (cached by the caller while the weights do not change).
    Returns (counts int32 [n_q], s_gt fp32 [n_q], scores or None).  `reduce_s_gt` is called on the
    target scores between the two launches (the cross-shard sum under catalogue sharding)."""
    if neg is not None:
        raise ValueError("the fused tensor-core path ranks against the full catalogue; use rank_from_scores for lists")
    W_hi, W_lo = W_split
    Q_hi, Q_lo = split_bf16(Q, passes == 3)
    n_q, d = Q.shape
    n = n1 - n0
    dev = Q.device
    gt = gt.contiguous()
    bias = _f(bias)
    ws = workspace.get(query("c2dsr_score_tc_workspace_bytes", n_q, n, d), dev)
    if s_gt is None:
        s_gt = torch.zeros(n_q, device=dev, dtype=F32)
        call("c2dsr_score_target_tc", ptr(Q_hi), ptr(Q_lo), ptr(W_hi), ptr(W_lo), ptr(bias), ptr(gt, I64), n_q, n0, n1,
             d, passes, ptr(n_q_limit), ptr(s_gt), ptr(ws), ws.numel(), stream())
        if reduce_s_gt is not None:
            reduce_s_gt(s_gt)
    if counts is None:
        counts = torch.zeros(n_q, device=dev, dtype=I32)
    lds = (n + 3) // 4 * 4
    S = torch.empty(n_q, lds, device=dev, dtype=F32) if want_scores else None
    call("c2dsr_score_count_tc", ptr(Q_hi), ptr(Q_lo), ptr(W_hi), ptr(W_lo), ptr(bias), ptr(s_gt), ptr(gt, I64), n_q,
         n0, n1, d, passes, ptr(n_q_limit), ptr(counts, I32), ptr(S), lds, None, 0, stream())
    return counts, s_gt, S
