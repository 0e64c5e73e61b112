"""Whole public path on one GPU: prepareLRInteraction + runLARIS(by_celltype=True, p-values) on a synthetic config.

    python tools/full_run.py CFG [rng] [n]

Prints wall-clock per call (host AnnData in, DataFrames out) - the "LARIS run in seconds" figure of the north star.
"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laris_b200 as la
from laris_b200.synth import make_dataset

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = sys.argv[2] if len(sys.argv) > 2 else "philox"
n = int(sys.argv[3]) if len(sys.argv) > 3 else None
t0 = time.perf_counter()
a, lr, c = make_dataset(cfg, n=n)
print(f"config {cfg}: n={c.n} G={c.G} P={c.P} C={c.C} N={c.N} rng={rng}  (synthesised in {time.perf_counter()-t0:.1f} s)", flush=True)
res = {}
for rep in range(2):                       # first pass warms the allocator / pinned pools
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lr_adata = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    laris_lr, ct = la.tl.runLARIS(lr_adata, a, n_nearest_neighbors=c.k, n_repeats=c.R, n_top_lr=c.P,
                                  number_nearest_neighbors=c.k, n_permutations=max(c.N, 100), rng=rng, verbose=False)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res = {"prepare_s": t1 - t0, "runLARIS_s": t2 - t1, "total_s": t2 - t0, "nnz_scores": int(lr_adata.X.nnz),
           "table_rows": int(len(ct))}
    print(f"pass {rep}: " + json.dumps(res), flush=True)
if os.environ.get("LARIS_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    lr_adata = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    la.tl.runLARIS(lr_adata, a, n_nearest_neighbors=c.k, n_repeats=c.R, n_top_lr=c.P, number_nearest_neighbors=c.k,
                   n_permutations=max(c.N, 100), rng=rng, verbose=False)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
