"""AdamW(amsgrad=True) as one fused CUDA launch over all live parameters.

Drop-in for ``torch.optim.AdamW(params, lr, weight_decay, amsgrad=True)`` as the reference
builds it (trainer.py:21-22): same ``param_groups`` (so ``StepLR`` works), same update
formulas, parameters whose ``.grad`` is None are skipped (the dead prototype-layer weights,
SURVEY.md Q3).  Gradients are read from ``p.grad`` -- which, as in the reference, keeps
accumulating across batches until ``zero_grad()`` is called once per epoch (Q2) -- or, for
data-parallel runs, from a caller-supplied all-reduced copy.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from ._cabi import AdamTensor, call, ptr, stream


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=True):
        if not amsgrad:
            raise ValueError("FusedAdamW implements the amsgrad variant only (the reference uses amsgrad=True)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=True))
        self._table_key = None
        self._table_dev = None
        self._table_host = None
        self.n_steps = 0          # bumped on every step(): parameters are updated through raw pointers

    @torch.no_grad()
    def step(self, closure=None, grads: Optional[Dict[torch.nn.Parameter, torch.Tensor]] = None):
        loss = closure() if closure is not None else None
        self.n_steps += 1
        for gi, group in enumerate(self.param_groups):
            live = []
            for p in group["params"]:
                g = grads.get(p) if grads is not None else p.grad
                if g is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                live.append((p, g, st))
            if not live:
                continue
            steps = {st["step"] for _, _, st in live}
            buckets = {s: [x for x in live if x[2]["step"] == s] for s in steps}   # normally a single bucket
            for s, items in buckets.items():
                self._launch(gi, group, items, s)
        return loss

    def _launch(self, gi, group, items, step_no):
        key = (gi, step_no == 0, tuple((p.data_ptr(), g.data_ptr()) for p, g, _ in items))
        dev = items[0][0].device
        if key != self._table_key:
            n = len(items)
            host = (AdamTensor * n)()
            for i, (p, g, st) in enumerate(items):
                if not (p.is_contiguous() and g.is_contiguous() and g.dtype == torch.float32):
                    raise RuntimeError("FusedAdamW needs contiguous fp32 parameters and gradients")
                host[i].p, host[i].g, host[i].acc = ptr(p), None, ptr(g)
                host[i].m, host[i].v, host[i].vmax = ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), ptr(st["max_exp_avg_sq"])
                host[i].n = p.numel()
            raw = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8)
            self._table_dev = raw.to(dev)
            self._table_key = key
            self._table_n = n
            self._table_max = max(p.numel() for p, _, _ in items)
        b1, b2 = group["betas"]
        call("c2dsr_adamw_amsgrad", ptr(self._table_dev), self._table_n, self._table_max, float(group["lr"]), b1, b2,
             group["eps"], group["weight_decay"], step_no, stream())
