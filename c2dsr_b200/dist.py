"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The path has exactly two exchange steps:

* training  -- data-parallel: the global batch is split across ranks, parameters and graph are
  replicated; per step one tiny all-reduce (valid-target counts and rows, so that loss normalisers
  equal the single-GPU ones), then a sharded optimiser step (``ShardedStep``): every rank updates
  1 / world of every tensor.  On NVLink-connected GPUs this runs over peer memory -- gradients and
  parameters of all ranks live in symmetric, peer-mapped blocks, and per tensor ONE kernel sums the
  gradient copies, applies AdamW to this rank's slice and stores the result into every rank's
  parameters (peer loads / stores, or multimem through the NVSwitch), bracketed by device-side
  barriers; otherwise per-tensor NCCL pipelines (reduce-scatter -> AdamW slice -> all-gather) and a
  flat bucket for the small tensors (``FlatShards``).  The optimiser moments of a tensor exist only
  as the slices their owners hold (``ShardedStep.big[i]["m" | "v" | "vmax" | "gsum"]``); like the
  reference, nothing is checkpointed;
* evaluation -- each rank encodes its slice of the query batch and the query vectors are
  all-gathered (2 MB); the item catalogue is sharded by rows of the classifier; the target score
  comes from the owning shard (sum all-reduce with zeros elsewhere, exact) and the per-shard partial
  rank counts are summed (int32 all-reduce, order independent, bit-exact).
"""
from __future__ import annotations

import os
import weakref
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
        # the data-parallel step runs ~10 streams side by side (branches, per-tensor optimiser pipelines with
        # cross-rank barriers): give every stream a hardware queue of its own (read when the CUDA context is made)
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
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

    def rebase(self, flat: torch.Tensor, bucket: torch.Tensor):
        """Move the flat parameter buffer and the gradient bucket into caller-supplied storage of the same size
        (the symmetric block of the peer-memory step)."""
        with torch.no_grad():
            flat.copy_(self.flat)
            bucket.zero_()
            self.flat, self.bucket = flat, bucket
            for p, off in zip(self.params, self.offsets):
                p.data = self.flat[off:off + p.numel()].view_as(p)
        self.views = [self.bucket[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]

    def stage_grads(self):
        """bucket <- this step's gradients."""
        # a parameter of the frozen layout that got no gradient this step (e.g. a classifier whose domain has no
        # valid row in this rank's shard) contributes zeros -- never a missing collective operand
        torch._foreach_copy_(self.views, [p.grad if p.grad is not None else torch.zeros_like(v)
                                          for p, v in zip(self.params, self.views)])

    def reduce_scatter_grads(self) -> torch.Tensor:
        """bucket <- this step's gradients; grad_shard <- sum over ranks of bucket[shard of this rank]."""
        self.stage_grads()
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
        early_ids = {id(p) for p in early}
        self.big = []
        small = []
        for p in params:
            if p.numel() >= self.BIG and p.numel() % (world_size * 4) == 0 and p.is_contiguous():
                per = p.numel() // world_size
                flat = p.data.view(-1)
                z = lambda: torch.zeros(per, dtype=torch.float32, device=p.device)
                st = dict(p=p, per=per, flat=flat, pshard=flat[rank * per:(rank + 1) * per], gshard=z(), m=z(), v=z(),
                          vmax=z(), gsum=z(), done=False, idx=len(self.big), early=id(p) in early_ids)
                host = (AdamTensor * 1)()
                host[0].p, host[0].g, host[0].acc = ptr(st["pshard"]), ptr(st["gshard"]), ptr(st["gsum"])
                host[0].m, host[0].v, host[0].vmax = ptr(st["m"]), ptr(st["v"]), ptr(st["vmax"])
                host[0].n = per
                st["table"] = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(p.device)
                self.big.append(st)
            else:
                small.append(p)
        self.small = FlatShards(small, rank, world_size) if small else None
        self.peer = self.small_peer = None
        if self.big and self.big[0]["p"].is_cuda and os.environ.get("C2DSR_DP", "peer") != "nccl":
            self._setup_peer()
        self.streams = tuple(torch.cuda.Stream() for _ in self.big)
        self._used = set()
        self.params = params
        early = {id(p) for p in early}
        self._hooks = []
        for st in self.big:
            if id(st["p"]) in early:
                self._hooks.append(st["p"].register_post_accumulate_grad_hook(lambda p, st=st: self._pipeline(st)))

    # ---- peer-memory path: gradient all-reduce + AdamW + parameter broadcast as ONE kernel over NVLink -------
    def _setup_peer(self):
        """Move the large tensors into one symmetric block per rank ([parameters | gradients], peer-mapped into
        every process and, where the NVSwitch offers it, bound to a multicast address) and build the tables of
        ``c2dsr_adamw_amsgrad_peer``.  The backward writes those gradients straight into the gradient half
        (``ops.GRAD_SINKS``).  Any failure to set this up -- no P2P mapping between the GPUs, a torch build without
        symmetric memory -- is agreed on by all ranks and leaves the NCCL pipeline in place."""
        from . import ops
        from ._cabi import PeerMap, PeerTensor, ptr
        dev = self.big[0]["p"].device
        n_small = self.small.shard * self.world_size if self.small is not None else 0
        total = sum(st["p"].numel() for st in self.big) + n_small
        ok = torch.ones(1, device=dev)
        block = hdl = None
        try:
            import torch.distributed._symmetric_memory as symm
            block = symm.empty(2 * total, dtype=torch.float32, device=dev)
            hdl = symm.rendezvous(block, dist.group.WORLD)
            if len(hdl.buffer_ptrs) != self.world_size:
                raise RuntimeError("symmetric memory: peer pointers missing")
        except Exception as e:                                   # noqa: BLE001 -- every failure means "use NCCL"
            ok.zero_()
            self.peer_error = repr(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            return
        block.zero_()
        # multicast (multimem.ld_reduce / multimem.st: the NVSwitch adds the copies and replicates the stores) moves
        # (1 + 1 / world) x the bytes per NVLink direction, plain peer loads / stores 2 (world - 1) / world x.
        # Measured on B200s (ms per step): 2 ranks 3.23 peer vs 3.97 multimem, 4 ranks 3.36 vs 3.47, 8 ranks 3.68 vs
        # 3.33 -- peer loads / stores (summed in rank order) up to 7 ranks, multicast from 8 on.
        # C2DSR_DP=p2p | multimem forces one of them.
        mode = os.environ.get("C2DSR_DP", "peer")
        want_mc = mode == "multimem" or (mode != "p2p" and self.world_size >= 8)
        mc = int(hdl.multicast_ptr) if want_mc and getattr(hdl, "has_multicast_support", False) else 0
        pmap = PeerMap()
        pmap.world, pmap.rank = self.world_size, self.rank
        for k in range(self.world_size):
            pmap.param[k] = int(hdl.buffer_ptrs[k])
            pmap.grad[k] = int(hdl.buffer_ptrs[k]) + 4 * total
        pmap.param_mc = mc or None
        pmap.grad_mc = (mc + 4 * total) if mc else None
        off = 0
        with torch.no_grad():
            for st in self.big:
                p, n, per = st["p"], st["p"].numel(), st["per"]
                view = block[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view
                st["flat"] = p.data.view(-1)
                st["pshard"] = st["flat"][self.rank * per:(self.rank + 1) * per]
                st["sink"] = block[total + off:total + off + n].view_as(p)
                ops.GRAD_SINKS[p.data_ptr()] = st["sink"]
                if not st["early"]:
                    me = weakref.ref(self)
                    ops.GRAD_READY[p.data_ptr()] = lambda st=st, me=me: me() is not None and me()._pipeline(st, staged=True)
                host = (PeerTensor * 1)()
                host[0].t.p, host[0].t.g, host[0].t.acc = ptr(st["pshard"]), None, ptr(st["gsum"])
                host[0].t.m, host[0].t.v, host[0].t.vmax = ptr(st["m"]), ptr(st["v"]), ptr(st["vmax"])
                host[0].t.n = per
                host[0].offset = off + self.rank * per
                st["peer_entry"] = bytes(host)
                st["peer_table"] = torch.frombuffer(bytearray(st["peer_entry"]), dtype=torch.uint8).to(dev)
                off += n
            if self.small is not None:
                # the small tensors' flat buffer and gradient bucket move into the block as one more "tensor"
                self.small.rebase(block[off:off + n_small], block[total + off:total + off + n_small])
                per = self.small.shard
                z = lambda: torch.zeros(per, dtype=torch.float32, device=dev)
                sp = dict(p=None, per=per, pshard=self.small.param_shard, gsum=z(), m=z(), v=z(), vmax=z(),
                          done=False, idx=len(self.big), early=False)
                host = (PeerTensor * 1)()
                host[0].t.p, host[0].t.g, host[0].t.acc = ptr(sp["pshard"]), None, ptr(sp["gsum"])
                host[0].t.m, host[0].t.v, host[0].t.vmax = ptr(sp["m"]), ptr(sp["v"]), ptr(sp["vmax"])
                host[0].t.n = per
                host[0].offset = off + self.rank * per
                sp["peer_entry"] = bytes(host)
                self.small_peer = sp
        self.peer = dict(block=block, hdl=hdl, map=pmap, total=total, multicast=bool(mc), tables={}, device=dev)
        weakref.finalize(self, ops.forget_grad_hooks, [st["p"].data_ptr() for st in self.big])
        hdl.barrier(channel=0)

    def _peer_barrier(self, channel: int):
        self.peer["hdl"].barrier(channel=channel, timeout_ms=60000)

    def _peer_stage(self, st):
        """The gradient of one tensor into the symmetric block (a no-op when the backward wrote it there)."""
        g = st["p"].grad
        if g is None:
            st["sink"].zero_()
        elif g.data_ptr() != st["sink"].data_ptr():
            st["sink"].copy_(g)

    def _peer_update(self, sts, stream, background: bool = False):
        """Gradients of ``sts`` complete on every rank -> one launch: all-reduce + AdamW + broadcast of the slices."""
        import ctypes as C
        from ._cabi import call, ptr
        key = tuple(st["idx"] for st in sts)
        table = self.peer["tables"].get(key)
        if table is None:
            raw = b"".join(st["peer_entry"] for st in sts)
            table = self.peer["tables"][key] = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(
                self.peer["device"])
        group = self.opt.param_groups[0]
        b1, b2 = group["betas"]
        call("c2dsr_adamw_amsgrad_peer", ptr(table), len(sts), max(st["per"] for st in sts), C.addressof(self.peer["map"]),
             ptr(self.opt.dyn_state), b1, b2, group["eps"], group["weight_decay"], int(background), stream.cuda_stream)

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
                 group["weight_decay"], 0, stream.cuda_stream)

    def _gather(self, st, stream):
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(st["flat"], st["pshard"])
        st["done"] = True

    def _pipeline(self, st, staged: bool = False):
        """The whole update of one tensor, started as soon as its gradient is complete: from the post-accumulate-
        grad hook of a classifier matrix, or (``staged``: the gradient already sits in the symmetric block) from
        the backward of an embedding table's branch."""
        if st["done"]:
            return
        stream = self._stream_of(st)
        if self.peer is not None:
            with torch.cuda.stream(stream):
                if not staged:
                    self._peer_stage(st)
                self._peer_barrier(1 + st["idx"])            # every rank's gradient of this tensor is in place
                # (short-lived "background" CTAs measured slower here: 3.36 vs 3.29 ms per step on 2 GPUs)
                self._peer_update([st], stream, background=os.environ.get("C2DSR_DP_BG", "0") != "0")
            st["done"] = True
            return
        self._reduce(st, stream)
        self._update(st, stream)
        self._gather(st, stream)

    def finish(self):
        """End of the backward: the tensors whose update has not started yet (the embedding tables), each on a
        stream of its own.  Kernels only overlap when SMs are free, and an AdamW launch fills the GPU: so ALL
        reduce-scatters are launched first (a collective kernel takes a few dozen CTAs), then the AdamW slices --
        each starts when its own reduce-scatter is done and runs beside the later ones -- then the all-gathers; last
        the small-tensor bucket on the caller's stream, and the join of the communication streams."""
        cur = torch.cuda.current_stream()
        if self.peer is not None:
            # the embedding tables, together, on the caller's stream: one barrier, one launch over their slices;
            # the small-tensor bucket (NCCL) meanwhile; one last barrier once every stream has joined -- after
            # it all parameter blocks are complete and all gradient blocks may be overwritten
            todo = [st for st in self.big if not st["done"]]
            for st in todo:
                self._peer_stage(st)
            if self.small is not None:
                self.small.stage_grads()
                todo.append(self.small_peer)
            if todo:
                self._peer_barrier(0)
                self._peer_update(todo, cur)
            self.opt.n_steps += 1
            if not torch.cuda.is_current_stream_capturing():
                self.opt.sync_lr()
        else:
            todo = [(st, self._stream_of(st)) for st in self.big if not st["done"]]
            for st, stream in todo:
                self._reduce(st, stream)
            for st, stream in todo:
                self._update(st, stream)
            for st, stream in todo:
                self._gather(st, stream)
        if self.peer is not None:
            pass
        elif self.small is not None:
            self.opt.step_flat(self.small.param_shard, self.small.reduce_scatter_grads())
            self.small.all_gather_params()
        else:
            self.opt.n_steps += 1
        for i in sorted(self._used):                  # (only streams that carry work of this step: capture-safe)
            cur.wait_stream(self.streams[i])
        self._used.clear()
        if self.peer is not None:
            self._peer_barrier(1 + len(self.big))
        for st in self.big:
            st["done"] = False
        for p in self.params:
            p.grad = None

    def close(self):
        """Forget the gradient sinks / hooks registered for this step's tensors (before the object is dropped)."""
        from . import ops
        for st in self.big:
            ops.GRAD_SINKS.pop(st["p"].data_ptr(), None)
            ops.GRAD_READY.pop(st["p"].data_ptr(), None)
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def zero_grad_sums(self):
        for st in self.big:
            st["gsum"].zero_()
        if self.small_peer is not None:
            self.small_peer["gsum"].zero_()
