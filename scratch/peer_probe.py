"""Data-parallel optimiser step in isolation (2+ GPUs): NCCL pipeline vs the peer-memory kernel (multimem / P2P).
Same gradients (rank-seeded) and parameters in every mode; prints agreement of the updated parameters between the
modes and across ranks, and the device time of ShardedStep.finish()."""
import json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from c2dsr_b200 import dist as cdist
from c2dsr_b200._cabi import call, ptr, query, stream
from c2dsr_b200.optim import FusedAdamW

rank, world, local = cdist.init_from_env("nccl")
dev = torch.device("cuda", local)
shapes = [(64094, 256)] * 3 + [(29208, 256), (34888, 256), (256, 256), (256,)]
res = {}
final = {}
for mode in ("nccl", "peer", "p2p"):
    os.environ["C2DSR_DP"] = mode
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.1) for s in shapes]
    opt = FusedAdamW(params, lr=1e-3, weight_decay=1e-2, amsgrad=True, accumulate=True)
    state = torch.zeros(query("c2dsr_step_state_bytes"), dtype=torch.uint8, device=dev)
    opt.attach_step_state(state)
    call("c2dsr_step_begin", ptr(state), 1, stream())
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)
    grads = [torch.randn(s, device=dev, generator=g) * 0.01 for s in shapes]
    for p, x in zip(params, grads):
        p.grad = x.clone()
    sh = cdist.ShardedStep(params, rank, world, opt, early=())
    opt.sharded = sh
    def one():
        for p, x in zip(params, grads):
            sink = next((st["sink"] for st in sh.big if st["p"] is p), None) if sh.peer is not None else None
            if sink is not None:
                sink.copy_(x); p.grad = sink
            else:
                p.grad = x
        sh.finish()
        call("c2dsr_step_begin", ptr(state), 1, stream())
    for _ in range(3):
        one()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(10):
        for p, x in zip(params, grads):
            p.grad = x
        if sh.peer is not None:
            for st in sh.big:
                st["sink"].copy_(grads[[id(q) for q in params].index(id(st["p"]))]); st["p"].grad = st["sink"]
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record(); sh.finish(); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        call("c2dsr_step_begin", ptr(state), 1, stream())
    t = torch.tensor(sorted(times)[len(times) // 2], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = {"finish_ms": round(float(t), 4), "peer": None if sh.peer is None else ("multimem" if sh.peer["multicast"] else "p2p"),
                 "peer_error": getattr(sh, "peer_error", None)}
    final[mode] = [p.detach().clone() for p in params]
    # all ranks hold the same parameters
    worst = 0.0
    for p in params:
        ref = p.detach().clone(); dist.broadcast(ref, 0)
        worst = max(worst, float((ref - p.detach()).abs().max()))
    res[mode]["max_diff_across_ranks"] = worst
for mode in ("peer", "p2p"):
    res[mode]["max_diff_vs_nccl"] = max(float((a - b).abs().max()) for a, b in zip(final[mode], final["nccl"]))
    res[mode]["max_abs_param"] = max(float(a.abs().max()) for a in final[mode])
if rank == 0:
    print(json.dumps({"world": world, **res}))
dist.destroy_process_group()
