"""Benchmark of the C2DSR hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1]: Food-Kitchen shape (29 207 + 34 886 items), d = 256, L = 15,
batch 256 per GPU, reference-default dropouts (0.2), synthetic logs.  A training "step" is the body of
the reference's epoch loop (trainer.py:47-49): convolve_graph() + train_batch() = 3 SpMM, 5 encoder
passes, infomax, 4 cross-entropies over the full catalogue, backward, AdamW-amsgrad.  The same run also
times full-catalogue evaluation (2 048 queries per batch) and, on rank 0 at N = 1, the CPU oracle port
on the host cores (`cpu_baseline`).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
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
    "tiny": dict(shape="tiny", d_latent=32, batch_size=32, batch_size_eval=64),      # contract tests only
}


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
    ap.add_argument("--breakdown", action="store_true", help="also print a per-entry-point time table to stderr")
    ap.add_argument("--score-path", type=str, default="tc", choices=["tc", "ffma"])
    ap.add_argument("--tc-passes", type=int, default=3, choices=[1, 3])
    ap.add_argument("--encoder-tc-passes", type=int, default=3, choices=[0, 1, 3],
                    help="encoder dense layers in training: 0 = fp32 FFMA, 3 = tcgen05 bf16 hi/lo split")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured training step")
    ap.add_argument("--all-rows", action="store_true",
                    help="run the loss GEMMs on every row, including those whose target is ignore_index")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def hyper(wl, dropout, device):
    from c2dsr_b200 import synth
    sh = synth.SHAPES[wl["shape"]]
    na, nb = sh["n_item_a"], sh["n_item_b"]
    return argparse.Namespace(
        data=wl["shape"], dataset=sh["dataset"], len_rec=10, n_neg_sample=999, d_latent=wl["d_latent"],
        shared_item_embed=False, d_bias=False, n_gnn=1, dropout_gnn=dropout, n_attn=1, n_head=1,
        dropout_attn=dropout, norm_first=False, lr=1e-3, l2=5e-4, lr_gamma=0.5, lr_step=10, len_max=sh["len_max"],
        lambda_loss=0.7, seed=3407, batch_size=wl["batch_size"], batch_size_eval=wl["batch_size_eval"],
        n_item_a=na, n_item_b=nb, n_item=na + nb + 1, idx_pad=na + nb, device=device, full_catalog=True,
        data_on_device=False, score_path="tc", tc_passes=3)


def make_workload(hp, n_train_batches, n_eval_batches, seed=0):
    """Synthetic logs -> graph from the whole training log, preprocessed tensors for the batches used."""
    import random
    from c2dsr_b200 import synth
    from c2dsr_b200.dataloader import preprocess_evaluate, preprocess_train
    from c2dsr_b200.graph import normalised_coo, transition_edges, _to_sparse
    sh = synth.SHAPES[hp.data]
    lengths = "full" if hp.data == "ee" else "fk"
    lm = hp.len_max - 1 if lengths == "full" else hp.len_max
    seqs = synth.make_sequences(sh["n_train"], hp.n_item_a, hp.n_item_b, len_max=lm, frac_a=sh["frac_a"], seed=seed,
                                lengths=lengths)
    shared, specific = transition_edges(seqs, hp.n_item_a)
    adj = (_to_sparse(normalised_coo(shared, hp.n_item), hp.n_item),
           _to_sparse(normalised_coo(specific, hp.n_item), hp.n_item))
    random.seed(seed)
    need = n_train_batches * hp.batch_size
    fields = preprocess_train(seqs[: int(need * 1.3) + 64], hp.n_item_a, hp.n_item_b, hp.len_max)
    while fields.shape[0] < need:
        fields = np.concatenate((fields, fields))
    fields = fields[:need]
    ev = synth.make_sequences(n_eval_batches * hp.batch_size_eval, hp.n_item_a, hp.n_item_b, len_max=lm,
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
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    timed.host_ms = (time.perf_counter() - t0) * 1e3 / max(n, 1)      # host time to enqueue one step
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


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
    n_tb = min(a.steps + a.warmup, 24)
    adj, fields, ev = make_workload(hp, n_tb * world, a.eval_batches, seed=0)
    fields = fields[rank::world] if world > 1 else fields          # each rank its own sequences
    host_tb, host_eb = train_batches(fields, B, True), eval_batches(ev, Bq, True)
    dev_tb = None        # device-resident batches come from the product's own loader (see below)
    dev_eb = [tuple(x.to(dev) for x in b) for b in host_eb]
    torch.manual_seed(hp.seed)
    ds = CDSRDataset.from_fields([fields[:, i] for i in range(14)], "train", L)
    hp.skip_ignored_rows = not a.all_rows
    loader = BatchLoader(ds, B, len_rec=R, ignore=(hp.n_item_a, hp.n_item_b))
    tr = Trainer.from_parts(hp, Quiet(), (loader, None, None), adj[0], adj[1])
    model = tr.model

    def train_step(batches):
        def f(i):
            return tr.train_step(batches[i % len(batches)])     # convolve_graph + train_batch (CUDA-graph replay)
        return f

    # ---- training, inputs resident in HBM: batches sliced on the device by the package's BatchLoader ----
    ds.to(dev)
    dev_tb = list(loader)
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

    # ---- full-catalogue evaluation ----
    model.eval()
    with torch.no_grad():
        model.convolve_graph()
        n_ev = max(a.eval_batches, 1) * 2
        ev_fn = lambda batches: (lambda i: tr.evaluate_batch(batches[i % len(batches)]))
        for i in range(2):
            ev_fn(dev_eb)(i)
        ev_dom = {"c2dsr_score_shard", "c2dsr_rank_from_scores", "c2dsr_pick_target", "c2dsr_score_count_tc",
                  "c2dsr_score_target_tc"}
        _cabi.PROFILE = {"names": ev_dom, "events": []}
        ms_ev = timed(ev_fn(dev_eb), n_ev, world)
        prof_ev, _cabi.PROFILE = _cabi.PROFILE, None
        ms_ev_e2e = timed(ev_fn(host_eb), n_ev, world)
    ev_dom_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in prof_ev["events"]) / n_ev
    eval_value = n_ev * Bq / (ms_ev / 1e3)
    eval_e2e = n_ev * Bq / (ms_ev_e2e / 1e3)

    # ---- optional per-entry breakdown (outside the timed regions) ----
    breakdown = None
    if a.breakdown:
        model.train()
        _cabi.PROFILE = {"names": None, "events": []}
        for i in range(3):
            model.convolve_graph()
            tr.train_batch(dev_tb[i % len(dev_tb)])
        torch.cuda.synchronize()
        agg = {}
        for nm, e0, e1 in _cabi.PROFILE["events"]:
            agg[nm] = agg.get(nm, 0.0) + e0.elapsed_time(e1) / 3
        _cabi.PROFILE = None
        breakdown = {k: round(v, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}

    pk = peaks()
    na, nb = hp.n_item_a, hp.n_item_b
    flops_ref = 3 * 2.0 * (2 * B * R) * d * (na + nb)               # fwd + 2x bwd over every row
    # rows that carry a target (the others are ignore_index: exactly zero loss and gradient, not computed)
    rows_frac = 1.0
    if not a.all_rows:
        used = [b.n_valid[0] for b in dev_tb if getattr(b, "n_valid", None)]
        if used:
            rows_frac = sum(ma * na + mb * nb for ma, mb in used) / (len(used) * (2.0 * B * R) * (na + nb))
    flops_step = flops_ref * rows_frac                               # work the path actually needs
    ach = flops_step / (dom_ms / 1e3) / 1e12
    ev_flops = 2.0 * Bq * d * (na * 0.5 + nb * 0.5)                 # per batch, domain mix ~50/50
    ev_ach = ev_flops / (ev_dom_ms / 1e3) / 1e12
    h2d = sum(x.numel() * x.element_size() for x in host_tb[0])
    out = {
        "metric": "train_seqs_per_sec", "value": round(train_value, 2), "unit": "seq/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2DSR {hp.dataset} shape, d={d}, L={L}, batch {B}/GPU, train step = convolve_graph"
                               " + train_batch (fwd, bwd, AdamW-amsgrad); + full-itemset eval",
                   "n_item_a": na, "n_item_b": nb, "len_rec": R, "dropout": a.dropout, "global_batch": B * world,
                   "parallelism": f"dp{world}", "cuda_graph_steps": bool(tr._graphs), "priming_steps": priming,
                   "gemm_arithmetic": {"classifier": f"{a.score_path} passes={a.tc_passes}",
                                       "encoder": f"passes={a.encoder_tc_passes}",
                                       "note": "passes=3: fp32 operands split into bf16 hi+lo, 3 tcgen05 MMAs per "
                                               "product, fp32 accumulate (product error ~1e-6); passes=0: fp32 FFMA"},
                   "host_enqueue_ms_per_step": round(host_ms, 3), "loss_rows": "all" if a.all_rows else
                   "rows with a target only (ignore_index rows contribute exactly 0 to loss and gradients)", "l2_note": "working set per step (params + AdamW state 1.3 GB, "
                   "logits 1.3 GB) exceeds the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": round(train_e2e, 2), "unit": "seq/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / a.steps, 4)},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "K4a classifier logits + cross-entropy (score_ce fwd+bwd, %d calls/step)" % n_dom_calls,
                     "bound": "tensor", "achieved": round(ach, 3), "peak": pk["tensor"], "unit": "TFLOP/s",
                     "frac": round(ach / pk["tensor"], 5),
                     # DRAM bytes of the K4a launches of one step, ncu --set full (profiles/r01_ncu_full_kernels.md)
                     # (profiles/r01_step_kernels_4p0ms.md: log-sum-exp, dZ-emitting, two gradient GEMMs, dZ column sums,
                     #  slab reductions, bf16 splits; rows with a target only)
                     "traffic": 3.29e9 if (a.score_path == "tc" and a.workload == "fk" and a.tc_passes == 3
                                           and not a.all_rows) else None,
                     "peak_source": pk["src"] + " sustained",
                     "algorithmic_gflop_per_step": round(flops_step / 1e9, 1),
                     "rows_with_target_frac": round(rows_frac, 4),
                     "reference_algorithm_gflop_per_step": round(flops_ref / 1e9, 1),
                     # tensor-core work actually issued: 4 GEMMs (forward, recompute, dH, dW) x passes MMAs per product
                     "executed_tflops": round(ach * (4.0 / 3.0) * a.tc_passes, 1) if a.score_path == "tc" else round(ach, 1),
                     "executed_frac": round(ach * (4.0 / 3.0) * a.tc_passes / pk["tensor"], 4) if a.score_path == "tc" else None,
                     "ms_per_step": round(dom_ms, 4), "share_of_step": round(dom_ms / (ms / a.steps), 4),
                     "path": ("tcgen05 bf16x%d: fused log-sum-exp forward, recompute + 2 gradient GEMMs backward" % a.tc_passes)
                     if a.score_path == "tc" else "ffma fp32 (materialised logits)"},
        "eval": {"metric": "full_catalog_eval_queries_per_sec", "value": round(eval_value, 1), "unit": "queries/s",
                 "batch": Bq, "batches": n_ev, "ms_per_batch": round(ms_ev / n_ev, 4),
                 "e2e": {"value": round(eval_e2e, 1), "unit": "queries/s",
                         "h2d_bytes_per_step": sum(x.numel() * x.element_size() for x in host_eb[0][:10]),
                         "d2h_bytes_per_step": Bq * 4},
                 "roofline": {"kernel": "K4b score + rank count", "bound": "tensor", "achieved": round(ev_ach, 3),
                              "peak": pk["tensor_burst"], "unit": "TFLOP/s", "frac": round(ev_ach / pk["tensor_burst"], 5),
                              "executed_tflops": round(ev_ach * a.tc_passes, 1) if a.score_path == "tc" else round(ev_ach, 1),
                              "ms_per_batch": round(ev_dom_ms, 4), "peak_source": pk["src"] + " burst",
                              "path": ("tcgen05 bf16x%d, fused count" % a.tc_passes) if a.score_path == "tc" else "ffma fp32"}},
    }
    if breakdown:
        out["breakdown_ms_per_step"] = breakdown
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(hp, adj, fields, ev, max_seconds=25.0)
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
def cpu_baseline(hp, adj, fields, ev, max_seconds=25.0, steps=None, warmup=1):
    """The oracle port (plain torch fp32, same algorithm as the reference) timed on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c2dsr_oracle as oracle
    torch.set_num_threads(os.cpu_count() or 1)
    h = {k: v for k, v in vars(hp).items() if isinstance(v, (int, float, bool, str))}
    otr = oracle.OracleTrainer(oracle.init_state(h, seed=1), adj[0].coalesce(), adj[1].coalesce(), h)
    B = hp.batch_size
    batches = [tuple(torch.from_numpy(np.ascontiguousarray(fields[i * B:(i + 1) * B, f])) for f in range(14))
               for i in range(min(4, fields.shape[0] // B))]
    otr.zero_grad()
    for i in range(warmup):
        otr.train_batch(batches[i % len(batches)], training=True)
    t0, n = time.perf_counter(), 0
    while True:
        otr.train_batch(batches[n % len(batches)], training=True)
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and (el > max_seconds * 0.6 or n >= 5)):
            break
    train_v = n * B / el
    six, four, neg = ev
    nq = 128
    eb = tuple(torch.from_numpy(np.ascontiguousarray(six[:nq, f])) for f in range(6)) + \
        tuple(torch.from_numpy(np.ascontiguousarray(four[:nq, f:f + 1])) for f in range(4)) + \
        (torch.from_numpy(np.ascontiguousarray(neg[:nq])),)
    otr.convolve_graph()
    t1 = time.perf_counter()
    otr.evaluate_batch(eb, full_catalog=True)
    ev_v = nq / (time.perf_counter() - t1)
    return {"value": round(train_v, 2), "unit": "seq/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} train steps of batch {B} after {warmup} warm-up ({el:.1f} s); eval {nq} queries",
            "eval_queries_per_sec": round(ev_v, 1), "host_cpus": os.cpu_count()}


def run_reference(a):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference is Python and cannot
    travel to the GPU box) on the host cores, same workload/metric.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = WORKLOADS[a.workload]
    hp = hyper(wl, a.dropout, torch.device("cpu"))
    adj, fields, ev = make_workload(hp, 4, 1, seed=0)
    n = max(1, min(a.steps, 6))
    w = max(0, min(a.warmup, 1))
    cb = cpu_baseline(hp, adj, fields, ev, steps=n, warmup=w)
    cb["sample"] = f"bounded: {n} of the requested {a.steps} steps, {w} warm-up; " + cb["sample"]
    out = {"impl": "reference", "metric": "train_seqs_per_sec", "value": cb["value"], "unit": "seq/s",
           "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(1e3 * hp.batch_size / cb["value"], 2),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"C2DSR {hp.dataset} shape, d={hp.d_latent}, L={hp.len_max}, batch {hp.batch_size}/GPU, "
                                  "train step = convolve_graph + train_batch (fwd, bwd, AdamW-amsgrad); + full-itemset eval",
                      "n_item_a": hp.n_item_a, "n_item_b": hp.n_item_b, "len_rec": hp.len_rec, "dropout": a.dropout,
                      "global_batch": hp.batch_size,
                      "implementation": "the reference's algorithm on the host CPU (oracle/c2dsr_oracle.py, plain "
                                        "torch ops, all host threads); bounded sample of the same workload"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "eval": {"metric": "full_catalog_eval_queries_per_sec", "value": cb["eval_queries_per_sec"],
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
