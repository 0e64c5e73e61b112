#!/usr/bin/env python
"""bench.py - headline benchmark of the LR scoring path (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic input:
kNN graph (incl. self) + decay weights + gene selection + fused diffusion/LR
score, producing the n x P score matrix (metric M1
"cell x LR-pair scores/s").  ``value`` times it with the inputs already in HBM, as
a CUDA-graph replay of the whole step (``engine.CapturedScoring``; the eager
schedule is timed beside it and reported as ``config.ms_per_step_eager``);
``e2e`` times ``laris_b200.tl.prepareLRInteraction`` (host AnnData in, host
AnnData with scipy CSR out: H2D, kernels, CSR compaction, D2H).  The
permutation metrics (random-graph null repeats/s, p-value permutations/s) and
the remaining stages are reported under ``extra``.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C]

N > 1 is launched by torchrun (one rank per GPU); LR pairs shard across ranks
with no data-path collective (weak scaling: every rank scores its own P pairs
over the same cells), timing is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell_x_lrpair_scores_per_sec"
UNIT = "scores/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--n", type=int, default=None, help="override the number of cells (debug)")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time the eager schedule instead of the CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(c):
    return (f"config{c.cfg_id}:{c.name} n={c.n} G={c.G} G_u<={c.G_u} P={c.P} k={c.k} C={c.C} R={c.R} N={c.N}")


def make_inputs(cfg, n, rank=0):
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(cfg, n=n)
    if rank:      # weak scaling: rank r scores its own P pairs over the same cells
        from laris_b200.synth import make_lr_table
        lr = make_lr_table(c, np.random.default_rng(5000 + 97 * rank + cfg))
    object.__setattr__(c, "cfg_id", cfg)
    return a, lr, c


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's prepareLRInteraction
# ---------------------------------------------------------------------------
def cpu_prepare(a, lr, k):
    """Reference laris/tools/__init__.py:80-117 as restated in oracle/restate.py
    (same sklearn / scipy routines the reference dispatches to; single thread,
    like the reference: scipy.sparse is serial and sklearn runs n_jobs=None)."""
    from oracle import restate as R
    t0 = time.perf_counter()
    li = R.gene_index(a.var_names, lr["ligand"])
    ri = R.gene_index(a.var_names, lr["receptor"])
    ind, dist = R.knn_graph(a.obsm["X_spatial"], k, True)
    S = R.lr_scores(a.X, ind, R.decay_weights(dist), li, ri)
    return time.perf_counter() - t0, S


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    a, lr, c = make_inputs(args.config, args.n)
    for _ in range(min(args.warmup, 1)):
        cpu_prepare(a, lr, c.k)
    times = [cpu_prepare(a, lr, c.k)[0] for _ in range(args.steps)]
    t = float(np.mean(times))
    val = c.n * c.P / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(c), "step": "prepareLRInteraction (graph + diffusion + LR scores)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"full workload, {args.steps} runs (reference path is single-threaded: "
                                   f"scipy.sparse serial, sklearn n_jobs=None; host has {os.cpu_count()} cores)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        """``indices``: GPU indices to sample in ONE nvidia-smi call per tick (rank 0 samples the whole job's GPUs;
        the other ranks pass an empty list and do not poll - eight pollers on one host perturb the run they watch)."""
        self.indices, self.rows, self.stop = list(indices), [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True) if self.indices else None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", ",".join(str(i) for i in self.indices)],
                                     capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1 if len(self.indices) == 1 else 0.25)

    def __enter__(self):
        if self.th is not None:
            self.th.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        if self.th is not None:
            self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")             # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)

    import laris_b200 as la
    from laris_b200 import _kernels as K, engine
    from laris_b200._lib import lib

    a, lr, c = make_inputs(args.config, args.n, rank)
    n, P, k = c.n, c.P, c.k
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")

    # ---- inputs resident in HBM ----
    xy_d = K.to_device(a.obsm["X_spatial"], torch.float64)
    csr_d = engine.upload_csr(a.X)
    lig_g, rec_g = la.tl._lr_gene_indices(a, lr)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)      # > 126 MB L2

    ev = lambda: torch.cuda.Event(enable_timing=True)
    state = {}

    # the LR table's plan (distinct genes, bank-conflict-free columns, packed pair indices) depends on
    # the table only; prepareLRInteraction caches it per table, the step reuses it the same way
    plan = engine.plan_lr_table(lig_g, rec_g, a.X.shape[1])
    layout = engine.choose_layout(a.X)

    def step(timed):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        graph = engine.build_graph(xy_d, k, True, None)
        xu = engine.select_planned(plan, csr_d, n, layout)
        e1.record()
        sm = engine.lr_scores(xu, graph, plan.lig_d, plan.rec_d)
        e2.record()
        state.update(graph=graph, xu=xu, sm=sm)
        return e0, e1, e2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The timed step is the CUDA-graph replay of the whole step (engine.CapturedScoring: the same ~20 kernels, no host
    # launches inside the timed region); the eager schedule is timed beside it.  The dominant kernel is timed through a
    # second, small graph holding only the fused score stage (pair-table pack + K2+K3).
    schedule = "cuda-graph replay (engine.CapturedScoring)"
    captured = kern_graph = None
    if not args.eager:
        try:
            captured = engine.CapturedScoring(xy_d, csr_d, plan, k, layout)
            kern_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(kern_graph):
                engine.lr_scores(captured.xu, captured.graph, plan.lig_d, plan.rec_d)
        except Exception as exc:                    # noqa: BLE001 - fall back to the eager schedule, say why
            captured = kern_graph = None
            schedule = f"eager (graph capture failed: {type(exc).__name__}: {str(exc)[:120]})"
            torch.cuda.synchronize()
    else:
        schedule = "eager (--eager)"

    def timed_region(run_step):
        out_step, out_kern = [], []
        for _ in range(args.steps):
            flush.fill_(1.0)                       # L2 flush between timed iterations (outside the event bracket)
            torch.cuda.synchronize()
            out = run_step()
            torch.cuda.synchronize()
            out_step.append(out[0].elapsed_time(out[-1]))
            if len(out) == 3:
                out_kern.append(out[1].elapsed_time(out[2]))
        return out_step, out_kern

    def replay_step():
        e0, e1 = ev(), ev()
        e0.record()
        captured.replay()
        e1.record()
        return e0, e1

    for _ in range(max(args.warmup, 3)):
        step(False)
        if captured is not None:
            captured.replay()
            kern_graph.replay()
    barrier()
    launches0 = lib.laris_launch_count()
    with ClockSampler(range(world) if rank == 0 else []) as clocks:
        barrier()
        wall0 = time.perf_counter()
        if captured is not None:
            step_ms, _ = timed_region(replay_step)
            launches = captured.launches * args.steps
        else:
            step_ms, kern_ms = timed_region(lambda: step(True))
            launches = lib.laris_launch_count() - launches0
        barrier()
        wall = time.perf_counter() - wall0
        if captured is not None:                   # beside it: the eager schedule, and the score stage alone
            eager_ms, _ = timed_region(lambda: step(True))
            kern_ms = []
            for _ in range(args.steps):
                flush.fill_(1.0)
                torch.cuda.synchronize()
                e0, e1 = ev(), ev()
                e0.record()
                kern_graph.replay()
                e1.record()
                torch.cuda.synchronize()
                kern_ms.append(e0.elapsed_time(e1))
            state.update(graph=captured.graph, xu=captured.xu, sm=captured.scores)
        else:
            eager_ms = step_ms

        # ---- e2e: the public API with host buffers (pinned inputs) ----
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
        import scipy.sparse as sp
        Xh = sp.csr_matrix((pin(a.X.data.astype(np.float32)), pin(a.X.indices.astype(np.int32)),
                            pin(a.X.indptr.astype(np.int64))), shape=a.X.shape)
        a_h = a.copy()
        a_h.X = Xh
        a_h.obsm["X_spatial"] = pin(a.obsm["X_spatial"])
        h2d = Xh.data.nbytes + Xh.indices.nbytes + Xh.indptr.nbytes + a_h.obsm["X_spatial"].nbytes
        out = None
        for _ in range(max(args.warmup, 3)):        # warm-up: page-locked staging buffers of both result generations
            out = la.tl.prepareLRInteraction(a_h, lr, number_nearest_neighbors=k)
        barrier()
        e2e_t = []
        for _ in range(args.steps):
            flush.fill_(1.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = la.tl.prepareLRInteraction(a_h, lr, number_nearest_neighbors=k)
            torch.cuda.synchronize()
            e2e_t.append(time.perf_counter() - t0)
        d2h = out.X.data.nbytes + out.X.indices.nbytes + out.X.indptr.nbytes
        barrier()

    t_step = float(np.sum(step_ms)) / 1e3
    t_e2e = float(np.sum(e2e_t))
    if world > 1:
        tt = torch.tensor([t_step, t_e2e, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_step, t_e2e, wall = tt.tolist()
    value = world * n * P * args.steps / t_step
    e2e_value = world * n * P * args.steps / t_e2e

    # ---- roofline of the dominant kernel (fused diffusion + LR score) ----
    # algorithmic bytes of K2+K3 per SURVEY 8(d): dense fp32 X_u read once, graph read once, dense fp32 S written once
    G_u = len(plan.genes)
    alg_bytes = 4 * n * G_u + 8 * n * k + 4 * n * P
    kname = "diffuse_dense_kernel" if layout == "dense" else "diffuse_score_kernel"
    k_ms = float(np.mean(kern_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                "bytes_per_score": alg_bytes / (n * P), "kernel_ms": k_ms, "layout": layout, "traffic": None}
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(f"config{args.config}", {}).get(kname)
        except Exception:
            pass

    extra = {}
    if not args.no_extra and rank == 0:
        extra = extra_metrics(torch, a, c, state, engine, K)
        if not args.no_cpu_baseline:
            extra.update(cpu_permutation_baselines(a, c, out.X, extra))

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        t_cpu, _ = cpu_prepare(a, lr, k)
        cpu = {"value": n * P / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"full workload once ({t_cpu:.2f} s); reference path is single-threaded "
                         f"(host has {os.cpu_count()} cores)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_step / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(c), "step": "kNN graph + decay weights + gene selection + fused "
                       "diffusion/LR score -> dense n x P score matrix in HBM",
                       "sharding": f"LR pairs: {P} per GPU x {world} GPUs, graph replicated, no collective",
                       "l2": "512 MiB buffer written between timed iterations (outside the event bracket)",
                       "schedule": schedule, "ms_per_step_eager": float(np.mean(eager_ms))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * t_e2e / args.steps, "ms_each": [round(1e3 * t, 3) for t in e2e_t],
                    "call": "laris_b200.tl.prepareLRInteraction(adata, lr_df) -> AnnData with scipy CSR on the host"},
            "gpu_launches": int(launches), "wall_s_timed_region": wall, "clocks": clocks.summary(),
            "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extra_metrics(torch, a, c, state, engine, K):
    """Permutation metrics (M2a/M2b) and the remaining stages, device-timed once each."""
    out = {}
    sm, n, P, k = state["sm"], c.n, c.P, c.k
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = ev(), ev()
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.min(ts)), r

    g2 = engine.build_graph(a.obsm["X_spatial"], k, False, 100.0)
    ms, _ = timed(lambda: K.spatial_stats(sm.S, P, g2.idx, g2.w, False, g2.order))
    out["observed_stats_ms"] = ms
    out["observed_stats_GBps_algorithmic"] = (4 * n * P + 8 * n * k) / ms / 1e6
    cols, wperm = K.null_graph_philox(n, k, g2.w, 0, 27)
    ms, _ = timed(lambda: K.spatial_stats(sm.S, P, cols, wperm, True, None))
    out["null_repeat_ms"] = ms
    out["null_repeats_per_sec"] = 1e3 / ms
    out["null_repeat_GBps_algorithmic"] = (4 * (1 + k) * n * P + 8 * n * k) / ms / 1e6
    labels = a.obs["CellTypes"].cat.codes.to_numpy().astype(np.int32)
    lab_d = K.to_device(labels, torch.int32)
    C = c.C
    N_ = K.neighborhood_composition(lab_d, g2.idx, g2.w, C)
    ms, _ = timed(lambda: K.celltype_pair_contract(sm.S, P, N_, 0, order=g2.order), reps=3)
    out["celltype_contract_ms"] = ms
    out["celltype_contract_impl"] = "tcgen05 (bf16 hi/mid/lo operand split, 6 MMAs per k-step, fp32 TMEM accumulation)"
    out["celltype_contract_TFLOPs_algorithmic"] = 2.0 * n * P * C * C / ms / 1e9
    if c.N > 0:
        rng = np.random.default_rng(0)
        table = K.to_device(rng.random((C * C, P)), torch.float64)
        bg = K.to_device(np.stack([rng.permutation(P)[:30] for _ in range(P)]).astype(np.int32), torch.int32)
        ms, _ = timed(lambda: K.pvalue_counts(table, bg, c.N, seed=27, only_ge=True))       # runLARIS default (standard p-value)
        out["pvalue_null_ms"] = ms
        out["pvalue_null_conditional_ms"] = timed(lambda: K.pvalue_counts(table, bg, c.N, seed=27))[0]
        out["pvalue_permutations_per_sec"] = c.N / (ms * 1e-3)
        out["pvalue_draws_per_sec"] = C * C * P * c.N / (ms * 1e-3)
    return out


def cpu_permutation_baselines(a, c, S_host, gpu):
    """The permutation half of the metric on the host cores (oracle port, 1 core - scipy.sparse / numpy are serial):
    one random-graph null repeat on the full score matrix, and the p-value resampling on a bounded sample of
    sender-receiver groups (its cost is linear in the number of groups)."""
    from oracle import restate as R
    import scipy.sparse as sp
    out = {}
    n, k, P, C = c.n, c.k, c.P, c.C
    S = sp.csr_matrix(S_host).astype(np.float64)
    ind, dist = R.knn_graph(a.obsm["X_spatial"], k, False)
    w = R.decay_weights(dist, 100.0)
    t0 = time.perf_counter()
    cols, wperm = next(iter(R.null_graphs(n, k, w, 1, 27, "numpy")))
    R.null_cosine(S, k, cols, wperm)
    t = time.perf_counter() - t0
    out["cpu_null_repeats_per_sec"] = 1.0 / t
    out["cpu_null_repeat_sample"] = f"1 repeat, full {n} x {P} score matrix ({t:.2f} s), oracle port, 1 core"
    if "null_repeats_per_sec" in gpu:
        out["null_repeats_speedup_vs_cpu"] = gpu["null_repeats_per_sec"] / out["cpu_null_repeats_per_sec"]
    if c.N > 0:
        rng = np.random.default_rng(0)
        groups = min(C * C, 40)                                    # bounded sample: 40 of the C^2 groups
        table = rng.random((groups, P))
        bg = np.stack([rng.permutation(P)[:30] for _ in range(P)]).astype(np.int64)
        rs = np.random.RandomState(27)
        t0 = time.perf_counter()
        R.pvalue_counts(table, np.ones(P, bool), bg, lambda g, pidx: rs.randint(0, 30, (len(pidx), c.N)))
        t = (time.perf_counter() - t0) * (C * C / groups)          # linear in the number of groups
        out["cpu_pvalue_permutations_per_sec"] = c.N / t
        out["cpu_pvalue_draws_per_sec"] = C * C * P * c.N / t
        out["cpu_pvalue_sample"] = f"{groups} of {C * C} groups x {P} pairs x {c.N} permutations, scaled linearly, oracle port, 1 core"
        if "pvalue_permutations_per_sec" in gpu:
            out["pvalue_permutations_speedup_vs_cpu"] = gpu["pvalue_permutations_per_sec"] / out["cpu_pvalue_permutations_per_sec"]
    return out


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
