"""``Trainer`` with the reference's interface (trainer.py:12-181): same constructor, attributes and
method names / return types, so ``main.py`` drives it unchanged.  The step itself runs on the
CUDA kernels behind the C ABI; torch supplies tensors, autograd bookkeeping and NCCL.

Differences that are deliberate and visible:
  * evaluation is batched (one device->host read per batch instead of one per sample) and can
    rank against the full catalogue (``args.full_catalog``) instead of ``list_neg``;
  * splits can live on the GPU (``args.data_on_device``, default on) so a batch is an index_select;
  * with ``torch.distributed`` initialised, training is data-parallel and evaluation shards the
    catalogue across ranks (dist.py).
"""
from __future__ import annotations

import time
from os.path import join

import torch

from . import _cabi, dist as cdist
from ._cabi import DynSeed, call, ptr, query, stream, workspace
from . import ops
from .c2dsr import C2DSR
from .dataloader import Batch, get_dataloader
from .graph import make_graph, make_graph_device
from .optim import FusedAdamW


class Trainer(object):
    def __init__(self, args, noter):
        self.rank, self.world_size = cdist.world()
        args.rank, args.world_size = self.rank, self.world_size
        self.trainloader, self.valloader, self.testloader = get_dataloader(args)
        if getattr(args, "device_graph", False) and getattr(args, "use_raw", False):
            self.adj_share, self.adj_specific = make_graph_device(args, join(args.path_raw, 'train_new.txt'))
        else:
            self.adj_share, self.adj_specific = make_graph(args, join(args.path_raw, 'train_new.txt'))
        self._setup(args, noter)

    @classmethod
    def from_parts(cls, args, noter, loaders, adj_share, adj_specific):
        """Build from already-made loaders and adjacencies (synthetic benchmarks, tests)."""
        self = cls.__new__(cls)
        self.rank, self.world_size = cdist.world()
        self.trainloader, self.valloader, self.testloader = loaders
        self.adj_share, self.adj_specific = adj_share, adj_specific
        self._setup(args, noter)
        return self

    def _setup(self, args, noter):
        self.args = args
        self.device = torch.device(args.device)
        self.model = C2DSR(args, self.adj_share, self.adj_specific).to(self.device)
        self.optimizer = FusedAdamW(filter(lambda x: x.requires_grad, self.model.parameters()), lr=args.lr,
                                    weight_decay=args.l2, amsgrad=True, accumulate=True)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=args.lr_step, gamma=args.lr_gamma)
        self.noter = noter
        for ld in (self.trainloader, self.valloader, self.testloader):
            if ld is not None and hasattr(ld.dataset, "check_bounds"):
                n_masked = ld.dataset.check_bounds(args.n_item)
                if n_masked and noter is not None and hasattr(noter, "log_msg"):
                    noter.log_msg(f"\t| note  | {ld.dataset.mode}: {n_masked} sequences start with a real item; their "
                                  "leading query rows have no allowed key (reference: NaN-prone, here: 0)")
        evals = [ld for ld in (self.valloader, self.testloader) if ld is not None and hasattr(ld.dataset, "pad_pos_zero")]
        self.model.pad_pos_zero = bool(getattr(args, "eval_pad_shortcut", True)) and len(evals) > 0 and \
            all(ld.dataset.pad_pos_zero for ld in evals)
        if getattr(args, "data_on_device", True):
            for ld in (self.trainloader, self.valloader, self.testloader):
                if ld is not None and hasattr(ld.dataset, "to"):
                    ld.dataset.to(self.device)
        self.n_tr = len(self.trainloader.dataset) if self.trainloader is not None else 0
        self.n_val = len(self.valloader.dataset) if self.valloader is not None else 0
        self.n_te = len(self.testloader.dataset) if self.testloader is not None else 0
        self.d_latent = args.d_latent
        self.n_item_a, self.n_item_b = args.n_item_a, args.n_item_b
        self.len_rec = args.len_rec
        self.lambda_loss = args.lambda_loss
        self.full_catalog = bool(getattr(args, "full_catalog", False))
        # "tc": tcgen05 GEMMs with the bf16 hi/lo split (3 passes = fp32-grade, 1 = plain bf16);
        # "ffma": fp32 CUDA-core GEMMs with materialised scores (the comparison path)
        self.score_path = getattr(args, "score_path", "tc")
        self.tc_passes = int(getattr(args, "tc_passes", 3))
        # skip loss rows whose target is ignore_index (exactly zero contribution); off = every row, like the reference
        self.skip_ignored = bool(getattr(args, "skip_ignored_rows", True))
        self._wsplit = {}
        self.shards = None           # data-parallel: flat parameter buffer + sharded optimiser step (lazy)
        # per-step device state: optimiser step number, learning rate and dropout key words (header:
        # c2dsr_step_state).  Eager and replayed steps read the same state, so they compute the same thing.
        self.step_state = torch.zeros(query("c2dsr_step_state_bytes"), dtype=torch.uint8, device=self.device)
        self.seed_base = (int(getattr(args, "seed", 0)) * 0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03 * (self.rank + 1)) \
            & 0xFFFFFFFFFFFFFFFF
        self.optimizer.attach_step_state(self.step_state)
        if self.world_size == 1 and self.device.type == "cuda" and bool(getattr(args, "early_adam", True)):
            # one GPU: the large tensors are updated beside the backward, as soon as their gradient is complete
            # (data parallel: dist.ShardedStep does the same with its per-tensor pipelines)
            m = self.model
            uniq = lambda ps: list({id(p): p for p in ps if p.requires_grad}.values())
            self.optimizer.enable_early(hooked=uniq([m.classifier_a.weight, m.classifier_b.weight]),
                                        staged=uniq([m.embed_i.weight, m.embed_i_a.weight, m.embed_i_b.weight]))
        call("c2dsr_step_begin", ptr(self.step_state), self.seed_base, stream())
        self.model.dyn_seed = DynSeed(self.step_state.data_ptr() + 8)
        # whole-step CUDA graphs; the data-parallel step captures its NCCL all-reduces too
        self.use_graph = bool(getattr(args, "cuda_graph", True)) and \
            (self.world_size == 1 or bool(getattr(args, "cuda_graph_dp", True)))
        self._graphs, self._warm, self._caps = {}, {}, None
        self._step_stream = None
        self._eval_graph, self._eval_seen = None, None

    def enable_pad_shortcut(self, eval_batches) -> bool:
        """For evaluation batches that do not come from this trainer's loaders (benchmarks, tests): check on the host
        that every PAD token has position 0 and, if so, let forward_select() use the pad-key shortcut."""
        pad = self.args.n_item - 1
        ok = all(bool((b[3 + k].cpu()[b[k].cpu() == pad] == 0).all()) for b in eval_batches for k in range(3))
        self.model.pad_pos_zero = ok and bool(getattr(self.args, "eval_pad_shortcut", True))
        return self.model.pad_pos_zero

    def _split_cache(self, weight, n0, n1):
        """bf16 (hi, lo) split of a classifier shard, refreshed whenever the weights change."""
        key = (weight.data_ptr(), n0, n1)
        ver = (weight._version, getattr(self.optimizer, "n_steps", 0))   # raw-pointer updates do not bump _version
        hit = self._wsplit.get(key)
        if hit is None:
            hit = [ver, ops.split_bf16(weight.detach()[n0:n1], self.tc_passes == 3)]
            self._wsplit[key] = hit
        elif hit[0] != ver:              # refreshed in place: captured evaluation graphs keep reading these buffers
            ops.split_bf16(weight.detach()[n0:n1], self.tc_passes == 3, out=hit[1])
            hit[0] = ver
        return hit[1]

    # ------------------------------------------------------------------------------------------
    def run_epoch(self):
        """trainer.py:40-71 -> (ranks_a, ranks_b) of the validation split."""
        self.model.train()
        self.optimizer.zero_grad()                      # once per epoch: gradients accumulate (Q2)
        sums = torch.zeros(3, device=self.device)
        t_start = time.time()
        for batch in self.trainloader:
            losses = self.train_step(batch)
            sums += torch.stack(losses).detach() * (getattr(batch, "global_batch", None)
                                                    or batch[0].shape[0] * self.world_size)
        loss_tr, loss_rec, loss_mi = (sums / max(self.n_tr, 1)).tolist()    # one host sync per epoch
        self.noter.log_train(loss_tr, loss_rec, loss_mi, time.time() - t_start)

        self.model.eval()
        pred_a, pred_b = [], []
        with torch.no_grad():
            self.model.convolve_graph()
            for ra, rb in self.evaluate_stream(self.valloader):
                pred_a += ra
                pred_b += rb
        return pred_a, pred_b

    def run_test(self):
        """trainer.py:73-83; reuses the propagated tables of the validation pass (Q13)."""
        self.model.eval()
        pred_a, pred_b = [], []
        with torch.no_grad():
            for ra, rb in self.evaluate_stream(self.testloader):
                pred_a += ra
                pred_b += rb
        return pred_a, pred_b

    def cal_mask(self, gt_mask):
        """trainer.py:85-89 (kept for API parity; the infomax kernel computes the weights itself)."""
        m = gt_mask.float()
        return (m / m.sum(-1, keepdim=True)).unsqueeze(-1).repeat(1, 1, self.d_latent)

    # ------------------------------------------------------------------------------------------
    def losses(self, batch):
        """Forward part of trainer.py:91-154 -> (loss, loss_rec, loss_mi), differentiable."""
        n_valid = self._caps if self._caps is not None else self._valid_rows(batch)
        # rows all ranks process in this step, known on the host (the loader's bookkeeping); equal shards otherwise
        g_rows = getattr(batch, "global_rows", None) or batch[0].shape[0] * self.world_size
        (seq_share, seq_a, seq_b, pos, pos_a, pos_b, gt_share_a, gt_share_b, gt_a, gt_b, gt_mask_a, gt_mask_b,
         seq_neg_a, seq_neg_b) = (x.to(self.device, non_blocking=True) for x in batch)
        m = self.model
        B, R, d = seq_share.shape[0], self.len_rec, self.d_latent
        h_share, hx, hy, h_neg_a, h_neg_b = m.forward_all(seq_share, seq_a, seq_b, pos, pos_a, pos_b, seq_neg_a,
                                                          seq_neg_b)
        # normalisers are global-batch quantities under data parallelism
        g_a, g_b = gt_a[:, -R:].reshape(-1), gt_b[:, -R:].reshape(-1)
        counts = torch.stack(((g_a != self.n_item_a).sum(), (g_b != self.n_item_b).sum(),
                              torch.full((), B, device=self.device))).float()     # (a fill: no host copy, no sync)
        cdist.allreduce_sum_(counts)
        n_a, n_b, b_glob = counts[0], counts[1], counts[2]

        loss_mi = ops.InfomaxFn.apply(h_share, hx, hy, h_neg_a, h_neg_b, m.D_a.weight, m.D_b.weight, m.D_a.bias,
                                      m.D_b.bias, gt_mask_a, gt_mask_b, 1.0 / g_rows)

        if self.score_path == "tc":
            # row assembly + tcgen05 logits / cross-entropy as one node per domain (ops.DomainLossTcFn)
            w_share = 1.0 / (R * b_glob)
            parts = []
            for k, (h_dom, cls, g_share, g_dom, n_dom) in enumerate(((hx, m.classifier_a, gt_share_a, gt_a, n_a),
                                                                     (hy, m.classifier_b, gt_share_b, gt_b, n_b))):
                rows = 2 * B * R if n_valid is None else min(int(n_valid[k]), 2 * B * R)
                parts.append(ops.DomainLossTcFn.apply(h_share, h_dom, cls.weight, cls.bias, m.classifier_pad.weight,
                                                      m.classifier_pad.bias, g_share, g_dom, w_share, n_dom, R, rows,
                                                      self.tc_passes))
            loss_rec = parts[0] + parts[1]
            loss = self.lambda_loss * loss_rec + (1 - self.lambda_loss) * loss_mi
            return loss, loss_rec, loss_mi

        hs = h_share[:, -R:].reshape(-1, d)
        ha, hb = hx[:, -R:].reshape(-1, d), hy[:, -R:].reshape(-1, d)
        share_w = (1.0 / (R * b_glob)).expand(B * R)
        parts = []
        for k, (h_dom, cls, g_share, g_dom, n_dom) in enumerate(((ha, m.classifier_a, gt_share_a, g_a, n_a),
                                                                 (hb, m.classifier_b, gt_share_b, g_b, n_b))):
            H = torch.cat((hs, hs + h_dom), 0)                 # item logits: h_share | h_share + h_dom
            Hpad = torch.cat((hs, h_dom), 0)                   # pad logit:   h_share | h_dom      (Q5)
            gt = torch.cat((g_share[:, -R:].reshape(-1), g_dom), 0)
            w = torch.cat((share_w, (1.0 / n_dom).expand(B * R)), 0)   # loss_share re-weighting (Q11)
            if n_valid is not None:
                # rows whose target is the ignore class add nothing to the loss or to any gradient: run the
                # catalogue-wide GEMMs on the other rows only (a stable partition brings them to the front)
                mv = n_valid[k]
                if mv == 0:                 # zero loss, and zero (not missing) gradients for the classifier
                    parts.append((hs.sum() + cls.weight.sum() + cls.bias.sum()) * 0.0)
                    continue
                keep = ops.compact_rows(gt, cls.weight.shape[0])[:mv]
                H, Hpad, gt, w = H.index_select(0, keep), Hpad.index_select(0, keep), gt[keep], w[keep]
            parts.append(ops.score_ce(H, Hpad, cls.weight, cls.bias, m.classifier_pad.weight, m.classifier_pad.bias,
                                      gt, w, path=self.score_path, passes=self.tc_passes))
        loss_rec = parts[0] + parts[1]
        loss = self.lambda_loss * loss_rec + (1 - self.lambda_loss) * loss_mi
        return loss, loss_rec, loss_mi

    def _valid_rows(self, batch):
        """(rows of the A-domain loss GEMM, rows of the B-domain one) that are not ignore_index, known on the
        host: from the loader's bookkeeping, or from the batch itself while it is still in host memory.  Returns
        None (= compute every row, as the reference does) when the option is off or the facts are on the device."""
        if not self.skip_ignored:
            return None
        nv = getattr(batch, "n_valid", None)
        if nv is not None and nv[1] == self.n_item_a and nv[2] == self.n_item_b:
            return nv[0]
        if not batch[6].is_cuda:
            R = self.len_rec
            f = [batch[i][:, -R:] for i in (6, 7, 8, 9)]
            return (int((f[0] != self.n_item_a).sum() + (f[2] != self.n_item_a).sum()),
                    int((f[1] != self.n_item_b).sum() + (f[3] != self.n_item_b).sum()))
        return None

    def train_batch(self, batch):
        """trainer.py:91-160 -> (loss, loss_rec, loss_mi) as 0-d tensors (global values under DP)."""
        loss, loss_rec, loss_mi = self.losses(batch)
        ops.ARMED[0] = True          # one gradient contribution per table / classifier: sinks + early optimiser steps
        try:
            loss.backward()
        finally:
            ops.ARMED[0] = False
        if self.world_size > 1:
            # sharded step: reduce-scatter the gradients, update this rank's 1/world of the flat parameter
            # buffer, all-gather the parameters (dist.FlatShards)
            if self.shards is None:
                live = [p for p in self.optimizer.param_groups[0]["params"] if p.grad is not None]
                m = self.model
                early = [] if getattr(self.args, "shared_item_embed", False) and False else \
                    [m.classifier_a.weight, m.classifier_b.weight]
                self.shards = cdist.ShardedStep(live, self.rank, self.world_size, self.optimizer, early=early)
                self.optimizer.sharded = self.shards
                # the parameters now live in the step's flat / symmetric buffers: nothing captured or cached
                # against their old storage may be replayed
                self._wsplit.clear()
                self._eval_graph = None
                self.model._pad_cache = {}
            self.shards.finish()
            call("c2dsr_step_begin", ptr(self.step_state), self.seed_base, stream())
            out = torch.stack((loss.detach(), loss_rec.detach(), loss_mi.detach()))
            cdist.allreduce_sum_(out)
            return out[0], out[1], out[2]
        self.optimizer.step()
        call("c2dsr_step_begin", ptr(self.step_state), self.seed_base, stream())       # state of the NEXT step
        return loss.detach(), loss_rec.detach(), loss_mi.detach()

    # ---- one iteration of the training loop (trainer.py:47-53), replayed from a CUDA graph ---------------
    def train_step(self, batch):
        """convolve_graph() + train_batch(batch).  The first two steps of a batch shape run eagerly (they
        create the .grad buffers, optimiser state and scratch the graph will reuse); the third is captured and
        every later one is a replay: ~300 kernel launches become one.  The loss GEMMs of the captured step run on
        a fixed row capacity (rows beyond the batch's valid count are ignore_index rows, which contribute
        nothing); a batch with more valid rows than that, or any other shape, takes the eager path."""
        if not (self.use_graph and self.model.training):
            self.model.convolve_graph(lazy=True)
            return self.train_batch(batch)
        nv = self._valid_rows(batch) if self.skip_ignored else None
        # (the infomax normaliser 1 / global rows is baked into the captured launch: part of the key)
        key = (tuple(batch[0].shape), nv is not None,
               getattr(batch, "global_rows", None) or batch[0].shape[0] * self.world_size)
        g = self._graphs.get(key)
        if g is None:
            seen = self._warm.setdefault(key, [])
            if len(seen) >= 2 and (nv is None or min(min(x) for x in seen) > 0):
                g = self._capture(key, batch, seen)
            else:
                seen.append(nv if nv is not None else (1, 1))
        if g is None or (nv is not None and (nv[0] > g["caps"][0] or nv[1] > g["caps"][1] or min(nv) == 0)):
            if g is not None:
                seen = self._warm.setdefault(key, [])
                seen.append(nv)
                g["overflows"] += 1
                if g["overflows"] >= 8:                                  # the capacity was a bad guess: re-capture
                    del self._graphs[key]
            self.model.convolve_graph(lazy=True)
            return self.train_batch(batch)
        packed = getattr(batch, "packed", None)
        if packed is not None and packed.shape == g["packed"].shape:
            g["packed"].copy_(packed, non_blocking=True)                 # one copy for the 14 fields
        else:
            for s, x in zip(g["static"], batch):
                s.copy_(x, non_blocking=True)
        self.optimizer.sync_lr()
        g["graph"].replay()
        self.model.hi_share, self.model.hi_a, self.model.hi_b = g["hi"]      # Q13: the step's propagations stay cached
        _cabi.REPLAYED_LAUNCHES += g["launches"]
        self.optimizer.n_steps += 1
        out = g["out"].clone()
        return out[0], out[1], out[2]

    def _capture(self, key, batch, seen):
        B, R = batch[0].shape[0], self.len_rec
        caps = None
        if key[1]:
            top = [max(x[k] for x in seen) for k in (0, 1)]
            caps = tuple(min(2 * B * R, -(-int(1.08 * t + 32) // 128) * 128) for t in top)
        if len({(tuple(x.shape), x.dtype) for x in batch}) == 1:
            packed_static = torch.empty((len(batch),) + tuple(batch[0].shape), dtype=batch[0].dtype, device=self.device)
            static = tuple(packed_static.unbind(0))
        else:
            packed_static = None
            static = tuple(torch.empty(x.shape, dtype=x.dtype, device=self.device) for x in batch)
        for s, x in zip(static, batch):
            s.copy_(x)
        self.optimizer.sync_lr()
        workspace.pinned = True            # scratch handed to a graph must never be freed by a later regrow
        # the cached propagations hold the previous step's autograd graph, and through it the AccumulateGrad
        # nodes of the embedding tables, which are tied to the (legacy default) stream they were created on:
        # drop them so that the capture builds its own
        self.model.hi_share = self.model.hi_a = self.model.hi_b = None
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = _cabi.launch_count()
        self._caps = caps
        # the step's own chain (and the model's branch streams) at high priority, the optimiser pipelines that run
        # beside the backward at the default, lowest one: their CTAs fill what the chain leaves free
        if self._step_stream is None:
            import os
            self._step_stream = torch.cuda.Stream(priority=-1 if os.environ.get("C2DSR_STEP_PRIO", "1") != "0" else 0)
        try:
            with torch.cuda.graph(graph, stream=self._step_stream):
                self.model.convolve_graph(lazy=True)
                step_in = Batch(static)             # carries the host-side normaliser facts of the captured shape
                step_in.global_rows = getattr(batch, "global_rows", None)
                step_in.global_batch = getattr(batch, "global_batch", None)
                out = torch.stack(self.train_batch(step_in))
        finally:
            self._caps = None
        g = dict(graph=graph, static=static, out=out, caps=caps or (2 * B * R, 2 * B * R),
                 launches=_cabi.launch_count() - l0, overflows=0,
                 packed=packed_static if packed_static is not None else torch.empty(0),
                 hi=(self.model.hi_share, self.model.hi_a, self.model.hi_b))
        self._graphs[key] = g
        return g

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def eval_queries(self, batch):
        """q_i = h_share[i, L-1] + (hx[i, idx_last_a] | hy[i, idx_last_b])  (trainer.py:169-177, Q8)."""
        seq_share, seq_a, seq_b, pos, pos_a, pos_b, idx_a, idx_b, xory, gt_last, list_neg = batch
        B_all = seq_share.shape[0]
        dom_b_all = xory.view(-1) != 0
        # the encoders are per-sequence work: each rank encodes its slice of the batch and the query
        # vectors are all-gathered (rows do not depend on which other rows share their launch)
        r0, r1 = cdist.shard_bounds(B_all, self.rank, self.world_size) if self.world_size > 1 else (0, B_all)
        seq_share, seq_a, seq_b, pos, pos_a, pos_b, idx_a, idx_b = (
            x[r0:r1] for x in (seq_share, seq_a, seq_b, pos, pos_a, pos_b, idx_a, idx_b))
        dom_b = dom_b_all[r0:r1]
        B = r1 - r0
        if B > 0:
            q = self._encode_queries((seq_share, seq_a, seq_b, pos, pos_a, pos_b, idx_a, idx_b, dom_b))
        else:
            q = torch.zeros(0, self.d_latent, device=self.device)
        return cdist.allgather_rows(q.contiguous(), B_all), dom_b_all

    def _encode_body(self, f):
        seq_share, seq_a, seq_b, pos, pos_a, pos_b, idx_a, idx_b, dom_b = f
        B, L = seq_share.shape
        # only one position per sequence and branch is read: h_share[:, L-1], hx[:, idx_last_a], hy[:, idx_last_b]
        last = torch.full((B,), L - 1, dtype=torch.int64, device=seq_share.device)
        hs, hx, hy = self.model.forward_select(seq_share, seq_a, seq_b, pos, pos_a, pos_b, last, idx_a.view(-1) % L,
                                               idx_b.view(-1) % L)
        return (hs + torch.where(dom_b.unsqueeze(-1), hy, hx)).contiguous()

    def _encode_queries(self, f):
        """The three encoders of an evaluation batch; replayed from a CUDA graph from the second batch of a shape
        on (the graph reads the cached propagations hi_*, so it is re-captured after every convolve_graph())."""
        m = self.model
        if not (self.use_graph and not m.training and m.hi_share is not None and f[0].is_cuda):
            return self._encode_body(f)
        key = (tuple(f[0].shape), m.hi_share.data_ptr(), m.hi_a.data_ptr(), m.hi_b.data_ptr())
        g = self._eval_graph
        if g is None or g["key"] != key:
            if self._eval_seen != key:                  # first batch after a convolve_graph / new shape: eager
                self._eval_seen = key
                return self._encode_body(f)
            static = tuple(x.clone() for x in f)
            workspace.pinned = True
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = _cabi.launch_count()
            with torch.cuda.graph(graph):
                q = self._encode_body(static)
            g = self._eval_graph = dict(key=key, graph=graph, static=static, q=q, launches=_cabi.launch_count() - l0)
        for s, x in zip(g["static"], f):
            s.copy_(x, non_blocking=True)
        g["graph"].replay()
        _cabi.REPLAYED_LAUNCHES += g["launches"]
        return g["q"]

    @torch.no_grad()
    def rank_queries(self, q, gt, neg, weight, bias):
        """rank = 1 + #{candidates : s > s_gt}; catalogue rows sharded across ranks when distributed."""
        n = weight.shape[0]
        n0, n1 = cdist.shard_bounds(n, self.rank, self.world_size)
        if neg is None and self.score_path == "tc" and q.shape[0] > 0:
            # full catalogue on tensor cores: fused GEMM + count, scores never reach HBM
            w_split = self._split_cache(weight, n0, n1)
            counts, _, _ = ops.score_rank_tc(q, w_split, bias[n0:n1], gt, n0, n1, passes=self.tc_passes,
                                             reduce_s_gt=cdist.allreduce_sum_)
            cdist.allreduce_sum_(counts)
            return counts + 1
        counts = torch.zeros(q.shape[0], dtype=torch.int32, device=q.device)
        s_gt = torch.zeros(q.shape[0], dtype=torch.float32, device=q.device)
        S = None
        if n1 > n0 and q.shape[0] > 0:
            S = ops.score_shard(q, weight[n0:n1], bias[n0:n1])
            s_gt = ops.pick_target(S, gt, n0, n1)
        cdist.allreduce_sum_(s_gt)
        if S is not None:
            ops.rank_from_scores(S, s_gt, gt, neg, n0, n1, counts)
        cdist.allreduce_sum_(counts)
        return counts + 1

    # ---- full-catalogue evaluation of a batch entirely on the device (one CUDA graph, one host read) ----------
    def _eval_fused_body(self, f, bufs=None):
        """f = the first ten evaluation fields on the device (global batch, identical on every rank).  Encoders on
        this rank's slice of the queries, all-gather of the query vectors, stable partition by domain on the device
        (c2dsr_eval_partition), per domain the target-score and counting GEMMs over this rank's catalogue shard with
        the domain's query count read from device memory, sum all-reduces of target scores and counts, ranks."""
        seq_share, seq_a, seq_b, pos, pos_a, pos_b, idx_a, idx_b, xory, gt_last = f
        Bg, d, dev = seq_share.shape[0], self.d_latent, self.device
        r0, r1 = cdist.shard_bounds(Bg, self.rank, self.world_size) if self.world_size > 1 else (0, Bg)
        dom, gt = xory.view(-1).contiguous(), gt_last.view(-1).contiguous()
        if r1 > r0:
            q = self._encode_body(tuple(x[r0:r1] for x in (seq_share, seq_a, seq_b, pos, pos_a, pos_b, idx_a, idx_b))
                                  + (dom[r0:r1] != 0,))
        else:
            q = torch.zeros(0, d, device=dev)
        q = cdist.allgather_rows(q.contiguous(), Bg)
        split = self.tc_passes == 3
        cap = -(-Bg // 128) * 128
        if bufs is None or bufs.get("cap") != (cap, d, split):
            # scratch of the ranking part; rows past a domain's query count keep stale (finite) values that are never
            # read back, so nothing but the counts needs clearing per batch
            bf = lambda: torch.zeros(cap, d, dtype=torch.bfloat16, device=dev)
            bufs = {} if bufs is None else bufs
            bufs.update(cap=(cap, d, split), QA_hi=bf(), QB_hi=bf(), QA_lo=bf() if split else None,
                        QB_lo=bf() if split else None, gts=torch.zeros(2, cap, dtype=torch.int64, device=dev),
                        slot=torch.zeros(Bg, dtype=torch.int32, device=dev),
                        n_ab=torch.zeros(2, dtype=torch.int32, device=dev),
                        s_gt=torch.zeros(2, cap, dtype=torch.float32, device=dev),
                        counts=torch.zeros(2, cap, dtype=torch.int32, device=dev),
                        out=torch.empty(2, Bg, dtype=torch.int32, device=dev))
        QA_hi, QB_hi, QA_lo, QB_lo = bufs["QA_hi"], bufs["QB_hi"], bufs["QA_lo"], bufs["QB_lo"]
        gts, slot, n_ab, s_gt, counts = bufs["gts"], bufs["slot"], bufs["n_ab"], bufs["s_gt"], bufs["counts"]
        call("c2dsr_eval_partition", ptr(q), ptr(dom), ptr(gt), Bg, d, ptr(QA_hi), ptr(QA_lo), ptr(QB_hi), ptr(QB_lo),
             ptr(gts[0]), ptr(gts[1]), ptr(slot), ptr(n_ab), stream())
        counts.zero_()
        for k, (cls, Q_hi, Q_lo) in enumerate(((self.model.classifier_a, QA_hi, QA_lo),
                                               (self.model.classifier_b, QB_hi, QB_lo))):
            # target scores of all queries from the replicated fp32 classifier: no exchange between the shards
            W, bias_full = cls.weight.detach(), cls.bias.detach()
            ws = workspace.get(query("c2dsr_score_tc_workspace_bytes", cap, W.shape[0], d), dev)
            call("c2dsr_score_target_full_tc", ptr(Q_hi), ptr(Q_lo), ptr(W), ptr(bias_full), ptr(gts[k]), cap, W.shape[0], d,
                 self.tc_passes, ptr(n_ab[k:]), ptr(s_gt[k]), ptr(ws), ws.numel(), stream())
            n0, n1 = cdist.shard_bounds(W.shape[0], self.rank, self.world_size)
            W_hi, W_lo = self._split_cache(cls.weight, n0, n1)
            bias = bias_full[n0:n1]
            call("c2dsr_score_count_tc", ptr(Q_hi), ptr(Q_lo), ptr(W_hi), ptr(W_lo), ptr(bias), ptr(s_gt[k]), ptr(gts[k]),
                 cap, n0, n1, d, self.tc_passes, ptr(n_ab[k:]), ptr(counts[k]), None, 0, None, 0, stream())
        cdist.allreduce_sum_(counts)                       # integer partial counts: order independent, bit-exact
        out = bufs["out"]
        call("c2dsr_eval_ranks", ptr(counts[0]), ptr(counts[1]), ptr(slot), ptr(dom), Bg, ptr(out), stream())
        return out

    def evaluate_stream(self, batches):
        """(rank_a, rank_b) for every batch of ``batches``, in order -- like evaluate_batch in a loop, but the host runs
        one batch ahead of the device: the inputs of batch i + 1 are copied and its graph is replayed before the ranks
        of batch i are read (pinned double buffer + event), so no device idle time is spent on the per-batch
        device -> host read that the reference's list-of-ints contract asks for."""
        pending = None
        for batch in batches:
            handle = self._evaluate_launch(batch)
            if pending is not None:
                yield self._evaluate_collect(pending)
            pending = handle
        if pending is not None:
            yield self._evaluate_collect(pending)

    @torch.no_grad()
    def _evaluate_launch(self, batch):
        fused = self.full_catalog and self.score_path == "tc" and batch[0].shape[0] > 0 and \
            self.model.attn_share.n_layers == 1 and self.model.hi_share is not None and self.use_graph
        if not fused:
            return ("done", self.evaluate_batch(batch))
        out = self._evaluate_fused(batch, raw=True)
        if not isinstance(out, torch.Tensor):
            return ("done", out)
        k = self._eval_slot = 1 - getattr(self, "_eval_slot", 0)
        bufs = self.__dict__.setdefault("_eval_pinned", {})
        key = (k, tuple(out.shape))
        if key not in bufs:
            bufs[key] = (torch.empty(out.shape, dtype=out.dtype, pin_memory=True), torch.cuda.Event())
        host, ev = bufs[key]
        host.copy_(out, non_blocking=True)
        ev.record()
        return ("wait", host, ev)

    @staticmethod
    def _evaluate_collect(handle):
        if handle[0] == "done":
            return handle[1]
        _, host, ev = handle
        ev.synchronize()
        r = host.numpy()
        return r[0][r[1] == 0].tolist(), r[0][r[1] != 0].tolist()

    def _evaluate_fused(self, batch, raw=False):
        m = self.model
        f = tuple(batch[:10])
        key = (tuple(f[0].shape), m.hi_share.data_ptr(), m.hi_a.data_ptr(), m.hi_b.data_ptr(), self.tc_passes)
        g = self._eval_graph
        if not self.use_graph or g is None or g["key"] != key:
            f_dev = tuple(x.to(self.device, non_blocking=True) for x in f)
            if not self.use_graph or self._eval_seen != key:       # first batch of a shape / phase: eager
                self._eval_seen = key
                out = self._eval_fused_body(f_dev, self.__dict__.setdefault("_eval_bufs", {}))
                r = out.cpu().numpy()
                return r[0][r[1] == 0].tolist(), r[0][r[1] != 0].tolist()
            static = tuple(x.clone() for x in f_dev)
            workspace.pinned = True
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = _cabi.launch_count()
            bufs = {}
            self._eval_fused_body(static, bufs)             # (allocates the scratch outside the capture)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                out = self._eval_fused_body(static, bufs)
            g = self._eval_graph = dict(key=key, graph=graph, static=static, out=out, bufs=bufs,
                                        keep=[dict(c) for c in m._pad_cache.values()],   # tensors the graph reads
                                        launches=(_cabi.launch_count() - l0) // 2)
        if all(x.is_cuda for x in f):
            torch._foreach_copy_(list(g["static"]), list(f))        # ten fields, one or two launches
        else:
            for s_, x in zip(g["static"], f):
                s_.copy_(x, non_blocking=True)
        g["graph"].replay()
        _cabi.REPLAYED_LAUNCHES += g["launches"]
        if raw:
            return g["out"]                                 # (evaluate_stream reads it one batch later)
        r = g["out"].cpu().numpy()                          # the batch's one device -> host read
        return r[0][r[1] == 0].tolist(), r[0][r[1] != 0].tolist()

    @torch.no_grad()
    def evaluate_batch(self, batch):
        """trainer.py:162-181 -> (rank_a, rank_b) Python lists in batch order."""
        if self.full_catalog and self.score_path == "tc" and batch[0].shape[0] > 0 and \
                self.model.attn_share.n_layers == 1 and self.model.hi_share is not None:
            return self._evaluate_fused(batch)
        xory_host = batch[8].view(-1).cpu()          # (before anything is enqueued: the device is idle or behind)
        # full-catalogue mode never reads list_neg: leave it on the host
        batch = tuple(x if (i == 10 and self.full_catalog) else x.to(self.device, non_blocking=True)
                      for i, x in enumerate(batch))
        gt_last, list_neg = batch[9].view(-1), batch[10]
        q, dom_b = self.eval_queries(batch)
        dom_b_host = xory_host != 0
        out = []
        for is_b, cls in ((False, self.model.classifier_a), (True, self.model.classifier_b)):
            sel = torch.nonzero(dom_b_host == is_b).view(-1).to(self.device)
            if sel.numel() == 0:
                out.append([])
                continue
            neg = None if self.full_catalog else list_neg.index_select(0, sel)
            ranks = self.rank_queries(q.index_select(0, sel).contiguous(), gt_last.index_select(0, sel), neg,
                                      cls.weight, cls.bias)
            out.append(ranks)
        return tuple(r.tolist() if isinstance(r, torch.Tensor) else r for r in out)
