"""world_size = 2 checks of the multi-GPU host logic on CPU (gloo): catalogue sharding with
partial rank counts, the sharded optimiser step, uneven data-parallel batches and the global loss normalisers.  The
per-rank compute is done by the oracle here (there is no GPU); the exchange logic under test is
the code the CUDA path uses (c2dsr_b200/dist.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c2dsr_oracle as oracle
    from c2dsr_b200 import dist as cdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    assert cdist.init_from_env("gloo")[:2] == (rank, world)
    assert cdist.world() == (rank, world)
    ok = True

    # --- query rows encoded per rank, all-gathered back in order (odd count: the last block is short) ---
    full = torch.arange(37 * 3, dtype=torch.float32).view(37, 3)
    r0, r1 = cdist.shard_bounds(37, rank, world)
    ok &= bool(torch.equal(cdist.allgather_rows(full[r0:r1].clone(), 37), full))

    # --- sharded optimiser step plumbing (FlatShards): parameters become views of one flat buffer, gradients are
    #     reduce-scattered, each rank updates its shard, the all-gather refreshes every parameter everywhere ---
    torch.manual_seed(7)
    ps = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2))]
    before = [p.detach().clone() for p in ps]
    sh = cdist.FlatShards(ps, rank, world)
    ok &= all(torch.equal(p.detach(), b) for p, b in zip(ps, before))            # values survive the re-pointing
    ok &= all(p.data_ptr() == sh.flat.data_ptr() + 4 * off for p, off in zip(ps, sh.offsets))
    for i, p in enumerate(ps):
        p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
    g = sh.reduce_scatter_grads()
    tot = sum(r + 1 for r in range(world))
    expect = torch.zeros(sh.shard * world)
    for i, (p, off) in enumerate(zip(ps, sh.offsets)):
        expect[off:off + p.numel()] = tot * (i + 1)
    ok &= bool(torch.equal(g, expect[rank * sh.shard:(rank + 1) * sh.shard]))
    sh.param_shard.sub_(0.5 * g)                                                   # "optimiser": plain SGD on the shard
    sh.all_gather_params()
    ok &= all(torch.allclose(p.detach(), b - 0.5 * tot * (i + 1)) for i, (p, b) in enumerate(zip(ps, before)))

    # --- catalogue-sharded ranking: identical on every rank, equal to the unsharded oracle ---
    rng = np.random.default_rng(0)
    n_q, N, d = 37, 1003, 16
    Q = rng.standard_normal((n_q, d)).astype(np.float32)
    W = rng.standard_normal((N, d)).astype(np.float32)
    gt = rng.integers(0, N, n_q)
    S_full = Q @ W.T
    ref = oracle.rank_from_scores(S_full, gt, None)
    n0, n1 = cdist.shard_bounds(N, rank, world)
    S = S_full[:, n0:n1]
    own = (gt >= n0) & (gt < n1)
    s_gt = torch.from_numpy(np.where(own, S_full[np.arange(n_q), gt], 0).astype(np.float32))
    cdist.allreduce_sum_(s_gt)
    ok &= bool(np.array_equal(s_gt.numpy(), S_full[np.arange(n_q), gt]))
    col = np.arange(n0, n1)[None, :]
    counts = torch.from_numpy(((S > s_gt.numpy()[:, None]) & (col != gt[:, None])).sum(1).astype(np.int32))
    cdist.allreduce_sum_(counts)
    ok &= bool(np.array_equal(counts.numpy() + 1, ref))

    # --- a parameter of the frozen layout that gets no gradient in some step contributes zeros (no rank may
    #     skip or fail a collective: the others would hang) ---
    ps[1].grad = None
    ps[0].grad = torch.full_like(ps[0], float(rank + 1))
    ps[2].grad = torch.zeros_like(ps[2])
    g2 = sh.reduce_scatter_grads()
    expect = torch.zeros(sh.shard * world)
    expect[sh.offsets[0]:sh.offsets[0] + ps[0].numel()] = tot
    ok &= bool(torch.equal(g2, expect[rank * sh.shard:(rank + 1) * sh.shard]))

    # --- BatchLoader under data parallelism: uneven last batch (9 samples on 2 ranks -> 5 + 4, normaliser 9) and a
    #     remainder shorter than the world (1 sample on 2 ranks -> replicated, normaliser 2): no empty shard ---
    from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
    for n, want_rows, want_glob, want_batch in ((29, (5, 4), 9, 9), (21, (1, 1), 2, 1)):
        fields = [torch.arange(n * 4).view(n, 4) + 1000 * f for f in range(14)]
        ds = CDSRDataset.from_fields(fields, "train", 4)
        batches = list(BatchLoader(ds, 10, rank=rank, world_size=world))
        ok &= len(batches) == 3 and batches[0][0].shape[0] == 5 and batches[0].global_rows == 10
        last = batches[-1]
        ok &= last[0].shape[0] == want_rows[rank] and last.global_rows == want_glob and last.global_batch == want_batch
        rows = torch.tensor([float(last[0].shape[0])])
        cdist.allreduce_sum_(rows)
        ok &= int(rows) == last.global_rows                    # the host-side normaliser equals the all-reduced one
        # every sample of the epoch is seen exactly once (replicated remainder: once per rank, weight 1 / world)
        seen = torch.cat([b[0][:, 0] for b in batches]).float()
        tot_seen = torch.tensor([seen.numel()], dtype=torch.float32)
        cdist.allreduce_sum_(tot_seen)
        ok &= int(tot_seen) == n + (want_glob - want_batch)

    # --- global normalisers: sum of per-rank valid counts ---
    c = torch.tensor([3.0 + rank, 7.0, 16.0])
    cdist.allreduce_sum_(c)
    ok &= c.tolist() == [7.0, 14.0, 32.0]
    ret[rank] = ok
    dist.destroy_process_group()


def test_two_rank_exchange_logic():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_shard_bounds_partition():
    from c2dsr_b200.dist import shard_bounds
    for n in (1, 7, 29207, 1000001):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
            assert all(hi >= lo for lo, hi in b)


def test_flat_shards_rebase_and_stage():
    """FlatShards.rebase moves the flat parameter buffer / gradient bucket into caller-supplied storage (the symmetric
    block of the peer-memory step): values preserved, parameters become views of the new buffer, padding stays zero;
    stage_grads copies the step's gradients into the bucket and writes zeros for a parameter without one."""
    from c2dsr_b200.dist import FlatShards
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s)) for s in ((5, 3), (7,), (2, 2, 2))]
    before = [p.detach().clone() for p in params]
    fs = FlatShards(params, rank=1, world_size=2)
    n = fs.flat.numel()
    assert n == 2 * fs.shard and fs.shard % FlatShards.ALIGN == 0
    block = torch.full((2 * n,), 7.0)
    fs.rebase(block[:n], block[n:])
    assert fs.flat.data_ptr() == block.data_ptr() and fs.bucket.data_ptr() == block[n:].data_ptr()
    for p, b, off in zip(params, before, fs.offsets):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == block[off:].data_ptr()                 # a view of the new buffer
    used = sum(p.numel() for p in params)
    assert float(fs.flat.abs().sum()) == float(sum(b.abs().sum() for b in before))   # padding copied as zeros
    assert float(block[n:].abs().sum()) == 0.0                                       # bucket cleared
    assert fs.param_shard.data_ptr() == block[fs.shard:].data_ptr() and fs.param_shard.numel() == fs.shard
    params[0].grad = torch.ones(5, 3)
    params[2].grad = torch.full((2, 2, 2), 2.0)                       # params[1] got no gradient this step
    fs.stage_grads()
    assert float(fs.bucket.sum()) == 15.0 + 16.0 and used == 30
    assert torch.equal(fs.views[1], torch.zeros(7))
    # an update through the flat buffer is an update of the parameters
    fs.flat.add_(1.0)
    assert torch.equal(params[1].detach(), before[1] + 1.0)
