"""Benchmark of the C2DSR hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload fk|mb|ee|1m|sweep-d256-L50|...] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE.json configs[1]: Food-Kitchen shape (29 207 + 34 886 items), d = 256, L = 15, batch 256
per GPU, reference-default dropouts (0.2), synthetic logs.  A training "step" is the body of the reference's epoch
loop (trainer.py:47-49): convolve_graph() + train_batch() = 3 SpMM, 5 encoder passes, infomax, 4 cross-entropies over
the full catalogue, backward, AdamW-amsgrad.  The same run also times full-catalogue evaluation (2 048 queries per GPU
and batch, catalogue sharded across the GPUs), the HBM-bound kernels on their own (`roofline_extra`) and, on rank 0
at N = 1, the reference itself on the host cores (`cpu_baseline`).  One JSON line is printed by rank 0.

Other workloads: `mb` (configs[2]), `ee` (configs[0] hyper-parameters), `1m` (configs[3]: evaluation only, 400 000 +
600 000 items), `sweep-d{128,256,512}-L{15,50,200}` (configs[4]).  `--impl reference` times the UNMODIFIED reference
(oracle/_ref, placed there by oracle/make_ref.py) on the host CPU with all cores, on the same synthetic workload.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "fk": dict(shape="fk", d_latent=256, batch_size=256, batch_size_eval=2048),
    "mb": dict(shape="mb", d_latent=256, batch_size=256, batch_size_eval=2048),
    "ee": dict(shape="ee", d_latent=128, batch_size=512, batch_size_eval=2048),
    "1m": dict(shape="1m", d_latent=256, batch_size=256, batch_size_eval=2048, eval_only=True),
    "tiny": dict(shape="tiny", d_latent=32, batch_size=32, batch_size_eval=64),      # contract tests only
}
for _d in (128, 256, 512):
    for _L in (15, 50, 200):
        WORKLOADS[f"sweep-d{_d}-L{_L}"] = dict(shape="fk", d_latent=_d, batch_size=256, batch_size_eval=2048, len_max=_L,
                                               lengths="uniform" if _L > 15 else "fk")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="fk", choices=sorted(WORKLOADS))
    ap.add_argument("--eval-batches", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.2)
    ap.add_argument("--breakdown", action="store_true", help="also print a per-entry-point time table in the JSON")
    ap.add_argument("--score-path", type=str, default="tc", choices=["tc", "ffma"])
    ap.add_argument("--tc-passes", type=int, default=3, choices=[1, 3])
    ap.add_argument("--encoder-tc-passes", type=int, default=3, choices=[0, 1, 3],
                    help="encoder dense layers in training: 0 = fp32 FFMA, 3 = tcgen05 bf16 hi/lo split")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured training step")
    ap.add_argument("--all-rows", action="store_true",
                    help="run the loss GEMMs on every row, including those whose target is ignore_index")
    ap.add_argument("--no-dp-parity", action="store_true", help="N > 1: skip the data-parallel parity check")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def hyper(wl, dropout, device):
    from c2dsr_b200 import synth
    sh = synth.SHAPES[wl["shape"]]
    na, nb = sh["n_item_a"], sh["n_item_b"]
    return argparse.Namespace(
        data=wl["shape"], dataset=sh["dataset"], len_rec=10, n_neg_sample=999, d_latent=wl["d_latent"],
        shared_item_embed=False, d_bias=False, n_gnn=1, dropout_gnn=dropout, n_attn=1, n_head=1,
        dropout_attn=dropout, norm_first=False, lr=1e-3, l2=5e-4, lr_gamma=0.5, lr_step=10,
        len_max=wl.get("len_max", sh["len_max"]), lambda_loss=0.7, seed=3407, batch_size=wl["batch_size"],
        batch_size_eval=wl["batch_size_eval"], n_item_a=na, n_item_b=nb, n_item=na + nb + 1, idx_pad=na + nb,
        device=device, full_catalog=True, data_on_device=False, score_path="tc", tc_passes=3,
        seq_lengths=wl.get("lengths", "full" if wl["shape"] == "ee" else "fk"), eval_only=bool(wl.get("eval_only")))


def common_config(hp, dropout, world):
    """The workload description both arms print (identical keys and values)."""
    what = "full-itemset eval only" if hp.eval_only else \
        "train step = convolve_graph + train_batch (fwd, bwd, AdamW-amsgrad); + full-itemset eval"
    return {"workload": f"C2DSR {hp.dataset} shape, d={hp.d_latent}, L={hp.len_max}, batch {hp.batch_size}/GPU, {what}",
            "n_item_a": hp.n_item_a, "n_item_b": hp.n_item_b, "d_latent": hp.d_latent, "len_max": hp.len_max,
            "len_rec": hp.len_rec, "dropout": dropout, "batch_per_gpu": hp.batch_size,
            "global_batch": hp.batch_size * world, "eval_batch_per_gpu": hp.batch_size_eval}


def make_workload(hp, n_train_batches, n_eval_batches, seed=0):
    """Synthetic logs -> graph from the whole training log, preprocessed tensors for the batches used."""
    import random
    from c2dsr_b200 import synth
    from c2dsr_b200.dataloader import preprocess_evaluate, preprocess_train
    from c2dsr_b200.graph import normalised_coo, transition_edges, _to_sparse
    sh = synth.SHAPES[hp.data]
    lengths = getattr(hp, "seq_lengths", "full" if hp.data == "ee" else "fk")
    lm = hp.len_max - 1 if lengths != "fk" else hp.len_max
    n_log = sh["n_train"] if sh["n_train"] else 200_000            # the 1M catalogue ships no training log: 200k sequences
    seqs = synth.make_sequences(n_log, hp.n_item_a, hp.n_item_b, len_max=lm, frac_a=sh["frac_a"], seed=seed,
                                lengths=lengths)
    shared, specific = transition_edges(seqs, hp.n_item_a)
    adj = (_to_sparse(normalised_coo(shared, hp.n_item), hp.n_item),
           _to_sparse(normalised_coo(specific, hp.n_item), hp.n_item))
    random.seed(seed)
    need = max(n_train_batches, 1) * hp.batch_size
    fields = preprocess_train(seqs[: int(need * 1.3) + 64], hp.n_item_a, hp.n_item_b, hp.len_max)
    while fields.shape[0] < need:
        fields = np.concatenate((fields, fields))
    fields = fields[:need]
    ev = synth.make_sequences(max(n_eval_batches, 1) * hp.batch_size_eval, hp.n_item_a, hp.n_item_b, len_max=lm,
                              frac_a=sh["frac_a"], seed=seed + 1, lengths=lengths)
    six, four, neg = preprocess_evaluate(ev, hp.n_item_a, hp.n_item_b, hp.len_max, hp.n_neg_sample)
    return adj, fields, (six, four, neg)


def train_batches(fields, B, pinned):
    out = []
    for i in range(fields.shape[0] // B):
        t = [torch.from_numpy(np.ascontiguousarray(fields[i * B:(i + 1) * B, f])) for f in range(14)]
        out.append(tuple(x.pin_memory() for x in t) if pinned else tuple(t))
    return out


def eval_batches(ev, Bq, pinned):
    six, four, neg = ev
    out = []
    for i in range(six.shape[0] // Bq):
        s = slice(i * Bq, (i + 1) * Bq)
        t = [torch.from_numpy(np.ascontiguousarray(six[s, f])) for f in range(6)] + \
            [torch.from_numpy(np.ascontiguousarray(four[s, f:f + 1])) for f in range(4)] + \
            [torch.from_numpy(np.ascontiguousarray(neg[s]))]
        out.append(tuple(x.pin_memory() for x in t) if pinned else tuple(t))
    return out


class Quiet:
    def log_train(self, *a):
        pass

    def log_msg(self, *a):
        pass


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while a timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        z = json.load(open(p))
        return dict(hbm=z["hbm_gbs"], tensor_burst=z["bf16_tflops"], tensor=z["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, src="fallback")


def max_over_ranks(ms, world):
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)
    return ms


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def timed(fn, n, world):
    """Exactly n calls of fn(i) between CUDA events; barrier + synchronize on both sides; max over ranks."""
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


def profiled(names, fn):
    """Run fn() with CUDA events around the named C-ABI entry points -> {entry: [ms, ...]}."""
    from c2dsr_b200 import _cabi
    _cabi.PROFILE = {"names": names, "events": []}
    try:
        fn()
        torch.cuda.synchronize()
        out = {}
        for nm, e0, e1 in _cabi.PROFILE["events"]:
            out.setdefault(nm, []).append(e0.elapsed_time(e1))
    finally:
        _cabi.PROFILE = None
    return out


# ------------------------------------------------------------------------------------------------
def dp_parity(rank, world, local_rank):
    """N > 1: the data-parallel path against the REFERENCE's golden losses (tests/golden/mid_default.npz, written by
    the reference itself on the full batches): each rank takes its slice of every batch, 8 steps through
    Trainer.train_step (eager, eager, capture with the NCCL collectives inside, replays)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import Golden, rel_err
    from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
    from c2dsr_b200.trainer import Trainer
    g = Golden("mid_default")
    hp = dict(g.hp)
    args = argparse.Namespace(**hp)
    args.device = torch.device("cuda", local_rank)
    args.encoder_tc_passes = 0
    n_steps = len(g.z["losses"])
    B = hp["batch_size"]
    train = CDSRDataset.from_fields([g.z["train_fields"][:n_steps * B, i] for i in range(14)], "train", hp["len_max"])
    loader = BatchLoader(train, B, rank=rank, world_size=world, len_rec=hp["len_rec"],
                         ignore=(hp["n_item_a"], hp["n_item_b"]))
    torch.manual_seed(hp["seed"])
    tr = Trainer.from_parts(args, Quiet(), (loader, None, None), g.adj("share"), g.adj("spec"))
    tr.model.load_state_dict({k: v.cuda() for k, v in g.group("init").items()})
    tr.model.train()
    tr.optimizer.zero_grad()
    worst = 0.0
    from c2dsr_b200 import dist as cdist
    big, cdist.ShardedStep.BIG = cdist.ShardedStep.BIG, 4096       # the fixture's tables take the large-tensor path
    try:
        for s, batch in enumerate(loader):
            got = [float(x) for x in tr.train_step(batch)]
            worst = max(worst, max(abs(a - b) / abs(b) for a, b in zip(got, g.z["losses"][s])))
    finally:
        cdist.ShardedStep.BIG = big
    path = "nccl" if tr.shards.peer is None else ("multimem" if tr.shards.peer["multicast"] else "p2p")
    n_big = len(tr.shards.big)
    final, d = g.group("final"), hp["d_latent"]
    w_err = 0.0
    for k, p in tr.model.state_dict().items():
        if k.endswith("attn_mask"):
            continue
        sl = slice(2 * d, None) if "in_proj" in k else slice(None)        # q / k rows: rounding-noise gradients (Q18)
        w_err = max(w_err, rel_err(p.cpu()[sl], final[k][sl]))
    out = {"fixture": "tests/golden/mid_default.npz (reference outputs)", "ranks": world, "steps": n_steps,
           "max_rel_loss_err_vs_reference": float(f"{worst:.3e}"), "max_rel_weight_err_vs_reference": float(f"{w_err:.3e}"),
           "graph_replayed": bool(tr._graphs), "dp_step": path, "large_tensors": n_big,
           "ok": bool(worst < 1e-4 and w_err < 1e-3)}
    tr._graphs.clear()
    tr.shards.close()
    del tr
    torch.cuda.synchronize()
    return out


# ------------------------------------------------------------------------------------------------
def run_b200(a):
    from c2dsr_b200 import _cabi, dist as cdist
    from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
    from c2dsr_b200.trainer import Trainer
    rank, world, local_rank = cdist.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    wl = WORKLOADS[a.workload]
    hp = hyper(wl, a.dropout, dev)
    hp.score_path, hp.tc_passes = a.score_path, a.tc_passes
    hp.encoder_tc_passes = a.encoder_tc_passes
    B, Bq, L, d, R = hp.batch_size, hp.batch_size_eval, hp.len_max, hp.d_latent, hp.len_rec
    pk = peaks()
    na, nb = hp.n_item_a, hp.n_item_b
    parity = None
    if world > 1 and not a.no_dp_parity and not hp.eval_only:
        parity = dp_parity(rank, world, local_rank)

    n_tb = 0 if hp.eval_only else min(a.steps + a.warmup, 24)
    # evaluation is weak-scaled too: every GPU brings 2 048 queries per batch, the catalogue is sharded
    adj, fields, ev = make_workload(hp, n_tb * world, a.eval_batches * world, seed=0)
    host_eb = eval_batches(ev, Bq * world, True)
    dev_eb = [tuple(x.to(dev) for x in b[:10]) + (b[10],) for b in host_eb]
    torch.manual_seed(hp.seed)
    hp.skip_ignored_rows = not a.all_rows
    out = {"metric": "train_seqs_per_sec", "value": None, "unit": "seq/s", "n_gpus": world, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": common_config(hp, a.dropout, world)}
    impl = {"parallelism": f"dp{world}", "gemm_arithmetic": {
        "classifier": f"{a.score_path} passes={a.tc_passes}", "encoder": f"passes={a.encoder_tc_passes}",
        "note": "passes=3: fp32 operands split into bf16 hi+lo, 3 tcgen05 MMAs per product, fp32 accumulate (product "
                "error ~1e-6; losses within 1e-4 of the reference, first-step gradients within 1e-3, see tests/); "
                "passes=0: fp32 FFMA"},
        "loss_rows": "all" if a.all_rows else "rows with a target only (ignore_index rows contribute exactly 0 to loss "
                                              "and gradients)",
        "l2_note": "working set per step (parameters + AdamW state 1.3 GB at FK) exceeds the 126 MB L2; no explicit flush"}

    if hp.eval_only:
        fields_r = np.zeros((B, 14, L), np.int64)
        ds = CDSRDataset.from_fields([fields_r[:, i] for i in range(14)], "train", L)
    else:
        fields_r = fields[rank::world] if world > 1 else fields          # each rank its own sequences
        ds = CDSRDataset.from_fields([fields_r[:, i] for i in range(14)], "train", L)
    loader = BatchLoader(ds, B, len_rec=R, ignore=(na, nb))
    tr = Trainer.from_parts(hp, Quiet(), (loader, None, None), adj[0], adj[1])
    model = tr.model
    impl["eval_pad_key_shortcut"] = tr.enable_pad_shortcut(host_eb)
    launches = 0
    extra = []

    if not hp.eval_only:
        host_tb = train_batches(fields_r, B, True)

        def train_step(batches):
            def f(i):
                return tr.train_step(batches[i % len(batches)])     # convolve_graph + train_batch (CUDA-graph replay)
            return f

        # ---- training, inputs resident in HBM: batches sliced on the device by the package's BatchLoader ----
        ds.to(dev)
        dev_tb = list(loader)
        for b in dev_tb:             # every rank feeds its own sequences: the global batch of a step is B x world
            b.global_rows = b.global_batch = B * world
        model.train()
        tr.optimizer.zero_grad()
        step = train_step(dev_tb)
        # CUDA events around the dominant entry points.  Set before the warm-up so that a step captured into a CUDA
        # graph carries them as event-record nodes (re-timed on every replay; read after the timed region = the
        # last timed step); eager steps append one pair per call.
        dom = {"c2dsr_score_ce_fwd", "c2dsr_score_ce_bwd", "c2dsr_score_ce_fwd_tc", "c2dsr_score_ce_bwd_tc"}
        tr.use_graph = tr.use_graph and not a.no_graph
        _cabi.PROFILE = {"names": dom, "events": [], "graph_events": []}
        # the step is captured on its third call (two eager calls create optimiser state and scratch first): when
        # fewer than three warm-up steps are asked for, untimed priming steps make up the difference
        priming = max(0, 3 - a.warmup)
        for i in range(priming + a.warmup):
            step(i)
        _cabi.PROFILE["events"].clear()
        l0 = _cabi.launch_count()
        clk = ClockSampler(local_rank)
        ms = timed(step, a.steps, world)
        clocks = clk.stop()
        # host time to enqueue a step, measured on 3 steps from an idle GPU (short enough not to fill the
        # driver's launch queue, so the host is never blocked by the device)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(3):
            step(i)
        host_ms = (time.perf_counter() - t0) * 1e3 / 3
        torch.cuda.synchronize()
        launches = _cabi.launch_count() - l0
        prof, _cabi.PROFILE = _cabi.PROFILE, None
        eager_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in prof["events"])
        gev = prof.get("graph_events") or []
        if gev:
            graph_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in gev)              # one replayed step
            eager_steps = len(prof["events"]) / max(len(gev), 1)
            dom_ms = (graph_ms * (a.steps - eager_steps) + eager_ms) / a.steps
            n_dom_calls = float(len(gev))
        else:
            dom_ms = eager_ms / a.steps
            n_dom_calls = len(prof["events"]) / a.steps
        train_value = a.steps * B * world / (ms / 1e3)

        # ---- training end to end: pinned host batches, H2D inside, loss read back every step ----
        step_h = train_step(host_tb)

        def e2e_step(i):
            float(step_h(i)[0])
        e2e_step(0)
        ms_e2e = timed(e2e_step, a.steps, world)
        train_e2e = a.steps * B * world / (ms_e2e / 1e3)

        flops_ref = 3 * 2.0 * (2 * B * R) * d * (na + nb)               # fwd + 2x bwd over every row
        rows_frac = 1.0
        if not a.all_rows:
            used = [b.n_valid[0] for b in dev_tb if getattr(b, "n_valid", None)]
            if used:
                rows_frac = sum(ma * na + mb * nb for ma, mb in used) / (len(used) * (2.0 * B * R) * (na + nb))
        flops_step = flops_ref * rows_frac                               # work the path actually needs
        ach = flops_step / (dom_ms / 1e3) / 1e12
        # logits-sized GEMMs issued per step for the 3 algorithmic ones: forward, then (d <= 256, fused backward) the
        # logits recomputed in each of the two backward kernels + dZ W + dZ^T H = 5; older path (d > 256): 4
        units = 5.0 if (a.score_path == "tc" and d <= 256) else 4.0
        out.update({"value": round(train_value, 2), "ms_per_step": round(ms / a.steps, 4), "clocks": clocks,
                    "e2e": {"value": round(train_e2e, 2), "unit": "seq/s",
                            "h2d_bytes_per_step": sum(x.numel() * x.element_size() for x in host_tb[0]),
                            "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / a.steps, 4)}})
        impl.update({"cuda_graph_steps": bool(tr._graphs), "priming_steps": priming,
                     "host_enqueue_ms_per_step": round(host_ms, 3)})
        if world > 1 and tr.shards is not None:
            peer = tr.shards.peer
            impl["dp_step"] = ("peer memory: all-reduce + AdamW + broadcast in one kernel (%s)" %
                               ("multimem over NVSwitch" if peer["multicast"] else "P2P loads / stores")) \
                if peer is not None else "NCCL reduce-scatter -> AdamW slice -> all-gather per tensor (%s)" % \
                getattr(tr.shards, "peer_error", "C2DSR_DP=nccl")
        tsrc = os.path.join("profiles", "r02_k4a_traffic.json")
        traffic = None
        if os.path.exists(os.path.join(ROOT, tsrc)) and a.workload == "fk" and a.score_path == "tc" and not a.all_rows:
            traffic = json.load(open(os.path.join(ROOT, tsrc))).get(f"passes{a.tc_passes}")
        out["roofline"] = {
            "kernel": "K4a classifier logits + cross-entropy (score_ce fwd+bwd, %d calls/step)" % n_dom_calls,
            "bound": "tensor", "achieved": round(ach, 3), "peak": pk["tensor"], "unit": "TFLOP/s",
            "frac": round(ach / pk["tensor"], 5), "traffic": traffic,
            "traffic_source": (tsrc + " (ncu dram__bytes_read+write of the K4a launches of one step, same command)")
            if traffic else None,
            "peak_source": pk["src"] + " sustained", "algorithmic_gflop_per_step": round(flops_step / 1e9, 1),
            "rows_with_target_frac": round(rows_frac, 4),
            "reference_algorithm_gflop_per_step": round(flops_ref / 1e9, 1),
            # tensor-core work actually issued: `units` logits-sized GEMMs (3 algorithmic) x passes MMAs per product
            "executed_tflops": round(ach * (units / 3.0) * a.tc_passes, 1) if a.score_path == "tc" else round(ach, 1),
            "executed_frac": round(ach * (units / 3.0) * a.tc_passes / pk["tensor"], 4) if a.score_path == "tc" else None,
            "ms_per_step": round(dom_ms, 4), "share_of_step": round(dom_ms / (ms / a.steps), 4),
            "path": ("tcgen05 bf16x%d: fused log-sum-exp forward; backward = two fused kernels (logits recomputed, dZ kept "
                     "in tensor memory, dH / dW accumulated in tensor memory)" % a.tc_passes)
            if a.score_path == "tc" else "ffma fp32 (materialised logits)"}

        # ---- the HBM-bound kernels on their own (CUDA events around the entry points, eager launches) ----
        nnz = [int(model.graph_share.nnz), int(model.graph_specific.nnz)]
        N_all = hp.n_item

        def spmm_bytes(nz):
            return nz * (8 + 4 * d) + 4 * (N_all + 1) + 2 * N_all * d * 4
        with torch.no_grad():
            model.eval()
            for _ in range(2):
                model.convolve_graph()
            ev_full = profiled({"c2dsr_spmm"}, lambda: [model.convolve_graph() for _ in range(3)])["c2dsr_spmm"]
            model.train()
            for _ in range(2):
                model.convolve_graph()
            tr_full = profiled({"c2dsr_spmm"}, lambda: [model.convolve_graph() for _ in range(3)])["c2dsr_spmm"]
        alg = (spmm_bytes(nnz[0]) + 2 * spmm_bytes(nnz[1])) / 3.0
        for name, ts in (("K2 SpMM, full product, no dropout (evaluation)", ev_full),
                         ("K2 SpMM, full product, input dropout %.1f (training semantics)" % a.dropout, tr_full)):
            t = statistics.median(ts)
            extra.append({"kernel": name, "bound": "hbm", "algorithmic_bytes": int(alg), "ms": round(t, 4),
                          "achieved": round(alg / t / 1e6, 1), "peak": pk["hbm"], "unit": "GB/s",
                          "frac": round(alg / t / 1e6 / pk["hbm"], 4)})
        model.train()
        tr.optimizer.early_enabled = False        # (timed on its own: the whole update at the end of the backward)
        per = profiled({"c2dsr_gather_fwd", "c2dsr_gather_bwd", "c2dsr_adamw_amsgrad_dyn", "c2dsr_adamw_amsgrad"},
                       lambda: [(model.convolve_graph(lazy=True), tr.train_batch(dev_tb[i % len(dev_tb)])) for i in range(3)])
        tr.optimizer.early_enabled = True
        T_step = 5 * B * L
        gf = sum(per.get("c2dsr_gather_fwd", [0.0])) / 3
        gb = sum(per.get("c2dsr_gather_bwd", [0.0])) / 3
        ad = (sum(per.get("c2dsr_adamw_amsgrad_dyn", [])) + sum(per.get("c2dsr_adamw_amsgrad", []))) / 3
        k1_bytes = 3 * T_step * d * 4 + T_step * 16
        n_par = sum(p.numel() for p in tr.optimizer.param_groups[0]["params"] if tr.optimizer.state.get(p))
        for name, byts, t in (("K1 gather forward (3 launches / step: 5 sequence tensors)", k1_bytes, gf),
                              ("K1 gather backward (deterministic scatter-add)", k1_bytes, gb),
                              ("AdamW-amsgrad + per-epoch gradient sum (48 B / parameter)", 48 * n_par, ad)):
            if t > 0:
                extra.append({"kernel": name, "bound": "hbm", "algorithmic_bytes": int(byts), "ms": round(t, 4),
                              "achieved": round(byts / t / 1e6, 1), "peak": pk["hbm"], "unit": "GB/s",
                              "frac": round(byts / t / 1e6 / pk["hbm"], 4)})

    # ---- full-catalogue evaluation: 2 048 queries per GPU and batch, catalogue sharded across the GPUs ----
    model.eval()
    with torch.no_grad():
        model.convolve_graph()
        n_ev = max(a.eval_batches, 1) * (4 if hp.eval_only else 2)
        if hp.eval_only:
            n_ev = max(n_ev, a.steps)
        ev_fn = lambda batches: (lambda i: tr.evaluate_batch(batches[i % len(batches)]))

        def ev_run(batches, n):
            # the trainer's own evaluation loop (run_epoch / run_test use it): ranks of every batch as Python lists,
            # the host one batch ahead of the device
            def once(_):
                for _ranks in tr.evaluate_stream(batches[i % len(batches)] for i in range(n)):
                    pass
            return once
        for i in range(max(3, a.warmup)):
            ev_fn(dev_eb)(i)
        l0 = _cabi.launch_count()
        clk = ClockSampler(local_rank)
        ms_ev = timed(ev_run(dev_eb, n_ev), 1, world)
        clocks_ev = clk.stop()
        ev_launches = _cabi.launch_count() - l0
        ms_ev_e2e = timed(ev_run(host_eb, n_ev), 1, world)
        # the ranking kernels alone: eager launches with events around the two entry points
        use_graph, tr.use_graph = tr.use_graph, False
        ev_prof = profiled({"c2dsr_score_count_tc", "c2dsr_score_target_tc", "c2dsr_score_shard",
                            "c2dsr_rank_from_scores", "c2dsr_pick_target"},
                           lambda: [ev_fn(dev_eb)(i) for i in range(3)])
        tr.use_graph = use_graph
    ev_dom_ms = sum(sum(v) for v in ev_prof.values()) / 3
    Bg = Bq * world
    eval_value = n_ev * Bg / (ms_ev / 1e3)
    eval_e2e = n_ev * Bg / (ms_ev_e2e / 1e3)
    dom_b = float(np.mean(ev[1][:, 2] != 0))
    # per GPU: all Bg queries against this GPU's shard of each catalogue (N / world items)
    ev_flops_gpu = 2.0 * Bg * d * (na * (1 - dom_b) + nb * dom_b) / world
    ev_ach = ev_flops_gpu / (ev_dom_ms / 1e3) / 1e12
    eval_obj = {"metric": "full_catalog_eval_queries_per_sec", "value": round(eval_value, 1), "unit": "queries/s",
                "batch_per_gpu": Bq, "global_batch": Bg, "batches": n_ev, "ms_per_batch": round(ms_ev / n_ev, 4),
                "e2e": {"value": round(eval_e2e, 1), "unit": "queries/s",
                        "h2d_bytes_per_step": sum(x.numel() * x.element_size() for x in host_eb[0][:10]),
                        "d2h_bytes_per_step": Bg * 8, "ms_per_batch": round(ms_ev_e2e / n_ev, 4)},
                "roofline": {"kernel": "K4b score + rank count (target + count GEMMs, both domains), per GPU",
                             "bound": "tensor", "achieved": round(ev_ach, 3), "peak": pk["tensor_burst"],
                             "unit": "TFLOP/s", "frac": round(ev_ach / pk["tensor_burst"], 5),
                             "flops_per_gpu_per_batch": ev_flops_gpu,
                             "executed_tflops": round(ev_ach * a.tc_passes, 1) if a.score_path == "tc" else round(ev_ach, 1),
                             "executed_frac": round(ev_ach * a.tc_passes / pk["tensor_burst"], 4) if a.score_path == "tc" else None,
                             "ms_per_batch": round(ev_dom_ms, 4), "share_of_batch": round(ev_dom_ms / (ms_ev / n_ev), 4),
                             "peak_source": pk["src"] + " burst",
                             "path": ("tcgen05 bf16x%d, fused count" % a.tc_passes) if a.score_path == "tc" else "ffma fp32"}}
    if hp.eval_only:
        out.update({"metric": "full_catalog_eval_queries_per_sec", "value": eval_obj["value"], "unit": "queries/s",
                    "steps": n_ev, "ms_per_step": eval_obj["ms_per_batch"], "clocks": clocks_ev, "e2e": eval_obj["e2e"],
                    "roofline": dict(eval_obj["roofline"], traffic=None)})
        launches = ev_launches
    out["gpu_launches"] = int(launches)
    out["eval"] = eval_obj
    if extra:
        out["roofline_extra"] = extra
    out["impl"] = impl
    if parity is not None:
        out["dp_parity"] = parity

    if a.breakdown and not hp.eval_only:
        model.train()
        agg = {}
        for nm, ts in profiled(None, lambda: [(model.convolve_graph(lazy=True), tr.train_batch(dev_tb[i % len(dev_tb)]))
                                              for i in range(3)]).items():
            agg[nm] = sum(ts) / 3
        out["breakdown_ms_per_step"] = {k: round(v, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(hp, adj, fields if not hp.eval_only else None, ev, max_seconds=20.0)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # captured graphs hold NCCL kernels: release them before the communicator goes away, and do not let a
        # stuck communicator teardown turn a finished measurement into a hang
        import gc
        tr._graphs.clear()
        tr._eval_graph = None
        gc.collect()
        torch.cuda.synchronize()
        t = threading.Thread(target=torch.distributed.destroy_process_group, daemon=True)
        t.start()
        t.join(timeout=20.0)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------------
def _plain(hp):
    return {k: v for k, v in vars(hp).items() if isinstance(v, (int, float, bool, str))}


def reference_run(hp, adj, fields, ev, steps, warmup, max_seconds=None, eval_queries=128):
    """The reference itself (oracle/_ref) if it is there, else the oracle port -- on the host CPU, all cores.
    -> (dict from the runner, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_run
    if fields is None:
        fields = np.zeros((hp.batch_size, 14, hp.len_max), np.int64)
    if ref_run.available():
        root = tempfile.mkdtemp(prefix="c2dsr_ref_")
        ref_run.write_processed(root, hp.dataset, hp.n_item_a, hp.n_item_b, fields[: 4 * hp.batch_size],
                                tuple(x[:max(eval_queries, 1)] for x in ev), adj)
        r = ref_run.time_reference(root, _plain(hp), fields, steps, warmup, eval_queries, max_seconds)
        return r, "reference"
    import c2dsr_oracle as oracle
    torch.set_num_threads(os.cpu_count() or 1)
    h = _plain(hp)
    otr = oracle.OracleTrainer(oracle.init_state(h, seed=1), adj[0].coalesce(), adj[1].coalesce(), h)
    B = hp.batch_size
    batches = [tuple(torch.from_numpy(np.ascontiguousarray(fields[i * B:(i + 1) * B, f])) for f in range(14))
               for i in range(max(1, min(4, fields.shape[0] // B)))]
    otr.zero_grad()
    for i in range(warmup):
        otr.train_batch(batches[i % len(batches)], training=True)
    t0, n = time.perf_counter(), 0
    while n < steps:
        otr.train_batch(batches[n % len(batches)], training=True)
        n += 1
        if max_seconds is not None and time.perf_counter() - t0 > max_seconds:
            break
    el = time.perf_counter() - t0
    six, four, neg = ev
    nq = eval_queries
    eb = tuple(torch.from_numpy(np.ascontiguousarray(six[:nq, f])) for f in range(6)) + \
        tuple(torch.from_numpy(np.ascontiguousarray(four[:nq, f:f + 1])) for f in range(4)) + \
        (torch.from_numpy(np.ascontiguousarray(neg[:nq])),)
    otr.convolve_graph()
    t1 = time.perf_counter()
    otr.evaluate_batch(eb, full_catalog=True)
    return dict(train_seq_per_s=n * B / el, seconds=el, steps=n, warmup=warmup, batch=B,
                eval_q_per_s=nq / (time.perf_counter() - t1), eval_queries=nq, threads=torch.get_num_threads()), "port"


def cpu_baseline(hp, adj, fields, ev, max_seconds=20.0):
    """`cpu_baseline` of the GPU arm: a bounded sample (about 20 s) of the same workload on the host cores."""
    steps = 0 if hp.eval_only else 1000
    r, kind = reference_run(hp, adj, fields, ev, steps=steps, warmup=0 if hp.eval_only else 1, max_seconds=max_seconds)
    what = "the unmodified reference (oracle/_ref: trainer.py Trainer, model.convolve_graph() + train_batch())" \
        if kind == "reference" else "oracle port (oracle/c2dsr_oracle.py; oracle/_ref absent)"
    if hp.eval_only:
        return {"value": round(r["eval_q_per_s"], 1), "unit": "queries/s", "cores": r["threads"], "kind": kind,
                "sample": f"{what}; {r['eval_queries']} evaluation queries (trainer.py:162-181)", "host_cpus": os.cpu_count()}
    return {"value": round(r["train_seq_per_s"], 2), "unit": "seq/s", "cores": r["threads"], "kind": kind,
            "sample": f"{what}; {r['steps']} train steps of batch {r['batch']} after {r['warmup']} warm-up "
                      f"({r['seconds']:.1f} s); eval {r['eval_queries']} queries",
            "eval_queries_per_sec": round(r["eval_q_per_s"], 1), "host_cpus": os.cpu_count()}


def run_reference(a):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) on the host cores, same workload, metric
    and unit.  Exactly --steps timed steps after --warmup warm-up steps; when that would take more than a few minutes
    at the full batch, each step is a bounded sample of the workload (a smaller batch, stated in `sample`).  Rank 0
    only; the other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = WORKLOADS[a.workload]
    world = max(a.gpus, 1)
    hp = hyper(wl, a.dropout, torch.device("cpu"))
    adj, fields, ev = make_workload(hp, 4, 1, seed=0)
    budget_s = 200.0
    note = ""
    if hp.eval_only:
        r, kind = reference_run(hp, adj, None, ev, steps=0, warmup=0, eval_queries=min(2048, max(128, a.steps * 16)))
        value, unit, ms_step = r["eval_q_per_s"], "queries/s", 1e3 * r["eval_queries"] / r["eval_q_per_s"] / max(a.steps, 1)
        sample = f"{r['eval_queries']} evaluation queries in total"
        steps_run = a.steps
    else:
        # one probe step at the full batch decides the per-step sample
        probe, kind = reference_run(hp, adj, fields, ev, steps=1, warmup=0, eval_queries=16)
        t_step = probe["seconds"]
        B_full = hp.batch_size
        if t_step * (a.steps + a.warmup) > budget_s:
            scale = budget_s / (t_step * (a.steps + a.warmup))
            hp.batch_size = int(max(16, min(B_full, B_full * scale // 16 * 16)))
            note = f"each step is a bounded sample: batch {hp.batch_size} of the workload's {B_full} "
        r, kind = reference_run(hp, adj, fields, ev, steps=a.steps, warmup=a.warmup)
        value, unit, ms_step = r["train_seq_per_s"], "seq/s", 1e3 * r["seconds"] / max(r["steps"], 1)
        sample = f"{note}{r['steps']} timed steps of batch {r['batch']} after {r['warmup']} warm-up ({r['seconds']:.1f} s); " \
                 f"eval {r['eval_queries']} queries"
        steps_run = r["steps"]
        hp.batch_size = B_full
    cb = {"value": round(value, 2), "unit": unit, "cores": r["threads"], "kind": kind, "sample": sample,
          "host_cpus": os.cpu_count()}
    out = {"impl": "reference", "metric": "full_catalog_eval_queries_per_sec" if hp.eval_only else "train_seqs_per_sec",
           "value": round(value, 2), "unit": unit, "n_gpus": a.gpus, "steps": steps_run, "warmup": a.warmup,
           "ms_per_step": round(ms_step, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": common_config(hp, a.dropout, world),
           "impl_config": {"implementation": "the unmodified reference from oracle/_ref (Trainer.train_batch + "
                                             "model.convolve_graph, stock PyTorch CPU ops, all host threads)"
                           if kind == "reference" else "oracle port (oracle/_ref absent)"},
           "cpu_baseline": cb,
           "e2e": {"value": round(value, 2), "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "eval": {"metric": "full_catalog_eval_queries_per_sec", "value": round(r["eval_q_per_s"], 1),
                    "unit": "queries/s"}}
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("C2DSR_HANG_DUMP"):          # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["C2DSR_HANG_DUMP"]), exit=True)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
