"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The path has exactly two exchange steps:

* training  -- data-parallel: the global batch is split across ranks, parameters and graph are
  replicated; per step one tiny all-reduce (valid-target counts and rows, so that loss normalisers
  equal the single-GPU ones), then a sharded optimiser step (``FlatShards``): the fp32 gradients are
  reduce-scattered, every rank runs AdamW on its 1/world of one flat parameter buffer, and the
  updated shards are all-gathered (same bytes on the wire as one all-reduce);
* evaluation -- each rank encodes its slice of the query batch and the query vectors are
  all-gathered (2 MB); the item catalogue is sharded by rows of the classifier; the target score
  comes from the owning shard (sum all-reduce with zeros elsewhere, exact) and the per-shard partial
  rank counts are summed (int32 all-reduce, order independent, bit-exact).
"""
from __future__ import annotations

import os
from typing import Tuple

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


class FlatShards:
    """Sharded optimiser step for data-parallel training (ZeRO-1 style; same bytes on the wire as one
    all-reduce, but every rank runs AdamW on 1/world of the parameters instead of all of them -- the update is
    HBM-bound, 0.56 ms for the 67 M parameters of the Food-Kitchen shape).

    The live parameters are re-pointed to views of ONE flat fp32 buffer (16-byte aligned offsets), which is cut
    into ``world`` equal contiguous shards.  Per step: gradients are copied into a flat bucket,
    reduce-scattered (sum) so that rank r holds the summed gradient of shard r, the optimiser updates
    ``flat[shard r]`` in place, and an in-place all-gather refreshes the other shards of ``flat`` -- i.e. of every
    parameter tensor -- on every rank.  Padding between tensors stays zero (AdamW maps 0 to 0)."""

    ALIGN = 4            # floats

    def __init__(self, params, rank: int, world_size: int):
        self.params = list(params)
        self.rank, self.world_size = rank, world_size
        dev = self.params[0].device
        self.offsets, o = [], 0
        for p in self.params:
            self.offsets.append(o)
            o += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        per = (o + world_size - 1) // world_size
        self.shard = (per + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        n = self.shard * world_size
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.bucket = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad_shard = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.views = [self.bucket[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]

    @property
    def param_shard(self) -> torch.Tensor:
        return self.flat[self.rank * self.shard:(self.rank + 1) * self.shard]

    def reduce_scatter_grads(self) -> torch.Tensor:
        """bucket <- this step's gradients; grad_shard <- sum over ranks of bucket[shard of this rank]."""
        # a parameter of the frozen layout that got no gradient this step (e.g. a classifier whose domain has no
        # valid row in this rank's shard) contributes zeros -- never a missing collective operand
        torch._foreach_copy_(self.views, [p.grad if p.grad is not None else torch.zeros_like(v)
                                          for p, v in zip(self.params, self.views)])
        if dist.get_backend() == "gloo":                     # CPU tests: gloo has no reduce-scatter
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM)
            self.grad_shard.copy_(self.bucket[self.rank * self.shard:(self.rank + 1) * self.shard])
        else:
            dist.reduce_scatter_tensor(self.grad_shard, self.bucket, op=dist.ReduceOp.SUM)
        return self.grad_shard

    def all_gather_params(self):
        """Every rank's updated shard -> the whole flat buffer (in place) on every rank."""
        if dist.get_backend() == "gloo":
            parts = [torch.empty_like(self.param_shard) for _ in range(self.world_size)]
            dist.all_gather(parts, self.param_shard.clone())
            self.flat.copy_(torch.cat(parts))
        else:
            dist.all_gather_into_tensor(self.flat, self.param_shard)


class ShardedStep:
    """Data-parallel optimiser step, version 2: every LARGE parameter tensor (embedding tables, classifier
    matrices: 98 % of the bytes) goes through its own pipeline on a communication stream --

        reduce-scatter of its gradient (straight from the tensor autograd produced: no 268 MB copy into a bucket)
        -> AdamW-amsgrad on this rank's 1 / world slice (with the slice of the per-epoch gradient sum, Q2)
        -> in-place all-gather of the updated slices into the parameter tensor

    -- and the pipelines of the classifier matrices start from a post-accumulate-grad hook, i.e. right after the
    K4a backward produced their gradients: their whole update is hidden behind the encoder / gather / GCN backward
    (nothing reads the classifier weights again in that step).  The embedding-table gradients only exist at the
    end of the backward; their three pipelines alternate between two streams so that the AdamW of one overlaps
    the collectives of the next.  The small tensors (encoder weights, biases: 2 %) share one flat bucket as
    before (``FlatShards``).  All of it is stream-ordered and is captured into the step's CUDA graph."""

    BIG = 1 << 20

    def __init__(self, params, rank: int, world_size: int, optimizer, early=()):
        from ._cabi import AdamTensor, ptr
        self.rank, self.world_size, self.opt = rank, world_size, optimizer
        params = list(params)
        self.big = []
        small = []
        for p in params:
            if p.numel() >= self.BIG and p.numel() % (world_size * 4) == 0 and p.is_contiguous():
                per = p.numel() // world_size
                flat = p.data.view(-1)
                z = lambda: torch.zeros(per, dtype=torch.float32, device=p.device)
                st = dict(p=p, per=per, flat=flat, pshard=flat[rank * per:(rank + 1) * per], gshard=z(), m=z(), v=z(),
                          vmax=z(), gsum=z(), done=False, idx=len(self.big))
                host = (AdamTensor * 1)()
                host[0].p, host[0].g, host[0].acc = ptr(st["pshard"]), ptr(st["gshard"]), ptr(st["gsum"])
                host[0].m, host[0].v, host[0].vmax = ptr(st["m"]), ptr(st["v"]), ptr(st["vmax"])
                host[0].n = per
                st["table"] = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(p.device)
                self.big.append(st)
            else:
                small.append(p)
        self.small = FlatShards(small, rank, world_size) if small else None
        self.streams = tuple(torch.cuda.Stream() for _ in self.big)
        self._used = set()
        self.params = params
        early = {id(p) for p in early}
        self._hooks = []
        for st in self.big:
            if id(st["p"]) in early:
                self._hooks.append(st["p"].register_post_accumulate_grad_hook(lambda p, st=st: self._pipeline(st)))

    def _stream_of(self, st):
        idx = st["idx"]
        self._used.add(idx)
        stream = self.streams[idx]
        stream.wait_stream(torch.cuda.current_stream())
        return stream

    def _reduce(self, st, stream):
        g = st["p"].grad
        with torch.cuda.stream(stream):
            if g is None:
                st["gshard"].zero_()
                dist.all_reduce(st["gshard"], op=dist.ReduceOp.SUM)       # (keeps the collective count equal on all ranks)
            else:
                dist.reduce_scatter_tensor(st["gshard"], g.contiguous().view(-1), op=dist.ReduceOp.SUM)

    def _update(self, st, stream):
        from ._cabi import call, ptr
        group = self.opt.param_groups[0]
        b1, b2 = group["betas"]
        with torch.cuda.stream(stream):
            call("c2dsr_adamw_amsgrad_dyn", ptr(st["table"]), 1, st["per"], ptr(self.opt.dyn_state), b1, b2, group["eps"],
                 group["weight_decay"], stream.cuda_stream)

    def _gather(self, st, stream):
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(st["flat"], st["pshard"])
        st["done"] = True

    def _pipeline(self, st):
        """The whole update of one tensor (post-accumulate-grad hook of the classifier matrices)."""
        if st["done"]:
            return
        stream = self._stream_of(st)
        self._reduce(st, stream)
        self._update(st, stream)
        self._gather(st, stream)

    def finish(self):
        """End of the backward: the tensors whose update has not started yet (the embedding tables), each on a
        stream of its own.  Kernels only overlap when SMs are free, and an AdamW launch fills the GPU: so ALL
        reduce-scatters are launched first (a collective kernel takes a few dozen CTAs), then the AdamW slices --
        each starts when its own reduce-scatter is done and runs beside the later ones -- then the all-gathers; last
        the small-tensor bucket on the caller's stream, and the join of the communication streams."""
        todo = [(st, self._stream_of(st)) for st in self.big if not st["done"]]
        for st, stream in todo:
            self._reduce(st, stream)
        for st, stream in todo:
            self._update(st, stream)
        for st, stream in todo:
            self._gather(st, stream)
        if self.small is not None:
            self.opt.step_flat(self.small.param_shard, self.small.reduce_scatter_grads())
            self.small.all_gather_params()
        else:
            self.opt.n_steps += 1
        cur = torch.cuda.current_stream()
        for i in sorted(self._used):                  # (only streams that carry work of this step: capture-safe)
            cur.wait_stream(self.streams[i])
        self._used.clear()
        for st in self.big:
            st["done"] = False
        for p in self.params:
            p.grad = None

    def zero_grad_sums(self):
        for st in self.big:
            st["gsum"].zero_()
