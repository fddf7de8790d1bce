"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The path has exactly two exchange steps:

* training  -- data-parallel: the global batch is split across ranks, parameters and graph are
  replicated; per step one tiny integer all-reduce (valid-target counts, so that loss
  normalisers equal the single-GPU ones) and one sum all-reduce of the fp32 gradients;
* evaluation -- each rank encodes its slice of the query batch and the query vectors are
  all-gathered (2 MB); the item catalogue is sharded by rows of the classifier; the target score
  comes from the owning shard (sum all-reduce with zeros elsewhere, exact) and the per-shard partial
  rank counts are summed (int32 all-reduce, order independent, bit-exact).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when no process group is initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """Initialise from torchrun's RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*.  Returns (rank, world, local_rank)."""
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", rank=rank, world_size=world_size,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world_size)
    return rank, world_size, local_rank


def shard_bounds(n: int, rank: int, world_size: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous shard [n0, n1) of ``n`` items for ``rank``; shards differ by at most ``align`` rows."""
    per = (n + world_size - 1) // world_size
    per = (per + align - 1) // align * align
    n0 = min(rank * per, n)
    return n0, min(n0 + per, n)


def allgather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Concatenate per-rank row blocks made with ``shard_bounds(n_total, rank, world)`` back into
    [n_total, ...] on every rank (one all-gather of equally padded blocks).  Identity for one process."""
    rank, world_size = world()
    if world_size == 1:
        return local
    per = (n_total + world_size - 1) // world_size
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((per * world_size,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return out[:n_total]


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks; identity for a single process."""
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class GradBucket:
    """Flat fp32 buffer holding a copy of every live gradient, all-reduced in one call.

    Each step's local gradients are summed over the ranks; the optimiser adds the reduced views to its
    per-epoch gradient sums (reference semantics: gradients accumulate across the batches of an epoch,
    trainer.py:42 vs :157; all-reduce is linear, so the order of the two sums does not matter).
    """

    def __init__(self):
        self.flat = None
        self.layout = None

    def reduce(self, params: Iterable[torch.nn.Parameter]) -> Dict[torch.nn.Parameter, torch.Tensor]:
        live = [p for p in params if p.grad is not None]
        layout = tuple((id(p), p.numel()) for p in live)
        total = sum(n for _, n in layout)
        if self.layout != layout:
            self.flat = torch.empty(total, dtype=torch.float32, device=live[0].device)
            self.layout = layout
        views, o = {}, 0
        for p in live:
            n = p.numel()
            v = self.flat[o:o + n].view_as(p)
            v.copy_(p.grad)
            views[p] = v
            o += n
        allreduce_sum_(self.flat)
        return views
