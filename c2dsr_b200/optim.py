"""AdamW(amsgrad=True) as one fused CUDA launch over all live parameters.

Drop-in for ``torch.optim.AdamW(params, lr, weight_decay, amsgrad=True)`` as the reference
builds it (trainer.py:21-22): same ``param_groups`` (so ``StepLR`` works), same update
formulas, parameters whose ``.grad`` is None are skipped (the dead prototype-layer weights,
SURVEY.md Q3).  Gradients are read from ``p.grad`` or, for data-parallel runs, from a
caller-supplied all-reduced copy.  As in the reference they accumulate across batches until
``zero_grad()`` is called once per epoch (Q2): either in ``p.grad`` itself (``accumulate=False``,
torch semantics) or in a sum owned by the optimiser (``accumulate=True``, see the class).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from ._cabi import AdamTensor, call, ptr, stream


class FusedAdamW(torch.optim.Optimizer):
    """``accumulate=True`` (what the Trainer uses): the optimiser owns the per-epoch gradient sum of Q2.  Each
    step the kernel does ``grad_sum += p.grad`` itself and updates from ``grad_sum``; ``p.grad`` is then
    released, so the next backward hands its gradient over without an add kernel per parameter (56 launches
    and ~1 GB of traffic per step at the Food-Kitchen shape).  ``zero_grad()`` clears the sums;
    ``accumulated_grad(p)`` returns what ``p.grad`` holds in the reference at the same point."""

    BIG = 1 << 20            # elements: tensors at least this large go into the first launch

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=True,
                 accumulate=False):
        if not amsgrad:
            raise ValueError("FusedAdamW implements the amsgrad variant only (the reference uses amsgrad=True)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=True))
        self.accumulate = bool(accumulate)
        self._table_key = None
        self._table_dev = None
        self._graph_tables = []   # pinned tables referenced by captured memcpy nodes: must outlive the graphs
        self.n_steps = 0          # bumped on every step(): parameters are updated through raw pointers
        self.early_enabled = True # (enable_early: set False to run every update at the end of the backward)
        # device-resident step state (c2dsr_step_state, see the header): when attached, the step number and the
        # learning rate are read by the kernel itself, which is what lets a whole training step be replayed
        # from a CUDA graph.  The caller advances it with c2dsr_step_begin after every step.
        self.dyn_state = None
        self._dyn_lr = None

    # ---- early steps: update a tensor the moment its gradient is complete, beside the rest of the backward ----
    def enable_early(self, hooked=(), staged=()):
        """The update is HBM-bound (48 B per parameter) while the encoder / gather backward that follows the K4a
        backward is latency-bound: the large tensors need not wait for the end of the backward.  ``hooked``:
        parameters whose single gradient contribution is complete at their post-accumulate-grad hook (classifier
        matrices); ``staged``: parameters whose producer announces the complete gradient itself (embedding tables:
        ``ops.GRAD_READY``, called on the producing stream).  Each gets a persistent gradient buffer
        (``ops.GRAD_SINKS``) the backward writes into, and its own stream; ``step()`` skips what is already done
        and joins the streams.  Needs the device step state; only armed around the Trainer's own backward
        (``ops.ARMED``)."""
        import weakref
        from . import ops
        self._early = {}
        self.early_enabled = True
        me = weakref.ref(self)
        for p in (*hooked, *staged):
            self._early[id(p)] = dict(p=p, sink=torch.zeros_like(p), stream=torch.cuda.Stream(), done=False, key=None,
                                      table=None)
            ops.GRAD_SINKS[p.data_ptr()] = self._early[id(p)]["sink"]
        self._early_hooks = [p.register_post_accumulate_grad_hook(
            lambda q, me=me: me() is not None and ops.ARMED[0] and me()._early_step(q)) for p in hooked]
        for p in staged:
            ops.GRAD_READY[p.data_ptr()] = lambda p=p, me=me: me() is not None and me()._early_step(p, staged=True)
        weakref.finalize(self, ops.forget_grad_hooks, [p.data_ptr() for p in (*hooked, *staged)])

    @torch.no_grad()
    def _early_step(self, p, staged: bool = False):
        e = self._early.get(id(p))
        if e is None or e["done"] or self.dyn_state is None or not self.early_enabled:
            return
        g = e["sink"]
        if not staged and (p.grad is None or p.grad.data_ptr() != g.data_ptr()):
            return                              # the gradient is not in the sink: the regular step takes it
        capturing = torch.cuda.is_current_stream_capturing()
        st = self.state[p]
        if len(st) == 0:
            if capturing:
                return
            st["step"] = 0
            for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq", "grad_sum"):
                st[k] = torch.zeros_like(p, memory_format=torch.preserve_format)
        key = (p.data_ptr(),) + tuple(st[k].data_ptr() for k in ("grad_sum", "exp_avg", "exp_avg_sq", "max_exp_avg_sq"))
        if e["key"] != key:
            if capturing:
                return
            host = (AdamTensor * 1)()
            host[0].p, host[0].g, host[0].acc = ptr(p), ptr(g), ptr(st["grad_sum"])
            host[0].m, host[0].v, host[0].vmax = ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), ptr(st["max_exp_avg_sq"])
            host[0].n = p.numel()
            e["table"] = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(p.device)
            e["key"] = key
        st["step"] += 1
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        if not capturing:
            self.sync_lr()
        side = e["stream"]
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            call("c2dsr_adamw_amsgrad_dyn", ptr(e["table"]), 1, p.numel(), ptr(self.dyn_state), b1, b2, group["eps"],
                 group["weight_decay"], int(os.environ.get("C2DSR_ADAM_BG", "1") != "0"), side.cuda_stream)
        e["done"] = True

    def attach_step_state(self, state: torch.Tensor):
        if len(self.param_groups) != 1:
            raise ValueError("the device step state carries one learning rate: use a single param group")
        self.dyn_state = state
        self._dyn_lr = float(self.param_groups[0]["lr"])
        call("c2dsr_step_state_set", ptr(state), self.n_steps, self._dyn_lr, stream())

    def sync_lr(self):
        """Push a learning rate changed by the scheduler to the device state (stream-ordered, no host sync)."""
        lr = float(self.param_groups[0]["lr"])
        if self.dyn_state is not None and lr != self._dyn_lr:
            call("c2dsr_step_state_set_lr", ptr(self.dyn_state), lr, stream())
            self._dyn_lr = lr

    # ---- sharded data-parallel step: one contiguous fp32 range of a flat parameter buffer (dist.FlatShards) ----
    @torch.no_grad()
    def step_flat(self, p_shard: torch.Tensor, g_shard: torch.Tensor):
        """AdamW-amsgrad on ``p_shard`` (in place) from the summed gradient ``g_shard`` of the same range; moments
        and the per-epoch gradient sum are kept for this range only.  Needs the device step state."""
        if self.dyn_state is None:
            raise RuntimeError("step_flat needs attach_step_state()")
        st = self.__dict__.get("_flat")
        if st is None or st["p"] != p_shard.data_ptr() or st["g"] != g_shard.data_ptr():
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the flat optimiser state must exist before a step is captured")
            z = lambda: torch.zeros_like(p_shard)
            st = self._flat = dict(p=p_shard.data_ptr(), g=g_shard.data_ptr(), m=z(), v=z(), vmax=z(), gsum=z())
            host = (AdamTensor * 1)()
            host[0].p, host[0].g, host[0].acc = ptr(p_shard), ptr(g_shard), ptr(st["gsum"])
            host[0].m, host[0].v, host[0].vmax = ptr(st["m"]), ptr(st["v"]), ptr(st["vmax"])
            host[0].n = p_shard.numel()
            st["table"] = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(p_shard.device)
        self.n_steps += 1
        group = self.param_groups[0]
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        b1, b2 = group["betas"]
        call("c2dsr_adamw_amsgrad_dyn", ptr(st["table"]), 1, p_shard.numel(), ptr(self.dyn_state), b1, b2, group["eps"],
             group["weight_decay"], 0, stream())

    def accumulated_grad(self, p):
        """The gradient sum since the last zero_grad() (= ``p.grad`` of the reference); None if p never had one."""
        if not self.accumulate:
            return p.grad
        return self.state[p].get("grad_sum") if p in self.state else None

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        if self.accumulate:
            sums = [st["grad_sum"] for st in self.state.values() if "grad_sum" in st]
            if sums:
                torch._foreach_zero_(sums)
        if self.__dict__.get("_flat") is not None:
            self._flat["gsum"].zero_()
        if self.__dict__.get("sharded") is not None:      # data parallel: per-tensor slices of the gradient sums
            self.sharded.zero_grad_sums()

    @torch.no_grad()
    def step(self, closure=None, grads: Optional[Dict[torch.nn.Parameter, torch.Tensor]] = None):
        loss = closure() if closure is not None else None
        self.n_steps += 1
        early = self.__dict__.get("_early") or {}
        done = [e for e in early.values() if e["done"]]
        for gi, group in enumerate(self.param_groups):
            live = []
            for p in group["params"]:
                e = early.get(id(p))
                if e is not None and e["done"]:
                    continue                    # updated already, beside the backward (enable_early)
                g = grads.get(p) if grads is not None else p.grad
                if g is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if self.accumulate:
                        st["grad_sum"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                live.append((p, g, st))
            if not live:
                continue
            steps = {st["step"] for _, _, st in live}
            buckets = {s: [x for x in live if x[2]["step"] == s] for s in steps}   # normally a single bucket
            for s, items in buckets.items():
                self._launch(gi, group, items, s)
            if self.accumulate:
                for p, _, _ in live:
                    p.grad = None               # consumed: the next backward's gradient is taken over as is
        cur = torch.cuda.current_stream() if done else None
        for e in done:
            cur.wait_stream(e["stream"])
            e["done"] = False
            if self.accumulate:
                e["p"].grad = None
        return loss

    def _launch(self, gi, group, items, step_no):
        # large tensors first: the table is launched in two parts (grid = chunks of the largest tensor x tensors),
        # so that the ~50 small tensors do not each pay for a grid sized for the embedding tables
        items = sorted(items, key=lambda it: -it[0].numel())
        key = (gi, step_no == 0, tuple((p.data_ptr(), g.data_ptr()) for p, g, _ in items))
        dev = items[0][0].device
        capturing = torch.cuda.is_current_stream_capturing()
        if key != self._table_key or capturing:
            n = len(items)
            host = (AdamTensor * n)()
            for i, (p, g, st) in enumerate(items):
                if not (p.is_contiguous() and g.is_contiguous() and g.dtype == torch.float32):
                    raise RuntimeError("FusedAdamW needs contiguous fp32 parameters and gradients")
                if self.accumulate:
                    host[i].p, host[i].g, host[i].acc = ptr(p), ptr(g), ptr(st["grad_sum"])
                else:
                    host[i].p, host[i].g, host[i].acc = ptr(p), None, ptr(g)
                host[i].m, host[i].v, host[i].vmax = ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), ptr(st["max_exp_avg_sq"])
                host[i].n = p.numel()
            raw = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8)
            # pinned staging + asynchronous copy: no host sync (torch's host allocator keeps the block until the
            # copy has run); inside a stream capture the copy becomes a memcpy node that re-reads the pinned
            # table on every replay, so that table is kept for the life of the process
            pinned = torch.empty(raw.numel(), dtype=torch.uint8, pin_memory=True)
            pinned.copy_(raw)
            if capturing:
                self._graph_tables.append(pinned)
            table = torch.empty(raw.numel(), dtype=torch.uint8, device=dev)
            table.copy_(pinned, non_blocking=True)
            self._table_dev = table
            self._table_key = None if capturing else key
            self._table_n = n
            n_big = sum(1 for p, _, _ in items if p.numel() >= self.BIG)
            self._table_parts = [(0, n_big, items[0][0].numel())] if n_big else []
            if n_big < n:
                self._table_parts.append((n_big, n - n_big, items[n_big][0].numel()))
        b1, b2 = group["betas"]
        if self.dyn_state is not None and not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        for first, count, max_n in self._table_parts:
            tab = ptr(self._table_dev) + first * C.sizeof(AdamTensor)
            if self.dyn_state is not None:
                call("c2dsr_adamw_amsgrad_dyn", tab, count, max_n, ptr(self.dyn_state), b1, b2, group["eps"],
                     group["weight_decay"], 0, stream())
            else:
                call("c2dsr_adamw_amsgrad", tab, count, max_n, float(group["lr"]), b1, b2, group["eps"],
                     group["weight_decay"], step_no, stream())
