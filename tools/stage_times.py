"""Per-stage device/wall timing of the hot path on one GPU (diagnostic; not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laris_b200 import _kernels as K, engine
from laris_b200.synth import make_dataset
import laris_b200 as la

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
a, lr, c = make_dataset(cfg, n=n)
k, P = c.k, c.P
xy = K.to_device(a.obsm["X_spatial"], torch.float64)
csr = engine.upload_csr(a.X)
lg, rg = la.tl._lr_gene_indices(a, lr)
genes, inv = np.unique(np.concatenate([lg, rg]), return_inverse=True)
layout = os.environ.get("LARIS_LAYOUT", "1") == "1"
cols, ncols = engine.lr_column_layout(inv[:P], inv[P:], len(genes)) if layout else (np.arange(len(genes), dtype=np.int32), len(genes))
gm = np.full(a.shape[1], -1, np.int32); gm[genes] = cols
gm_d = K.to_device(gm, torch.int32)
lig = K.to_device(cols[inv[:P]].astype(np.int32), torch.int32); rec = K.to_device(cols[inv[P:]].astype(np.int32), torch.int32)
labels = K.to_device(a.obs["CellTypes"].cat.codes.to_numpy().astype(np.int32), torch.int32)

def T(name, fn, reps=5):
    reps = int(os.environ.get("LARIS_REPS", reps))
    ev, wl = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); wl.append(time.perf_counter() - t0)
        ev.append(e0.elapsed_time(e1))
    print(f"{name:34s} dev min {min(ev):9.3f} ms  med {np.median(ev):9.3f} ms | wall med {1e3*np.median(wl):9.3f} ms", flush=True)
    return r

print(f"config {cfg}: n={c.n} G_u={len(genes)} P={P} k={k} C={c.C}")
idx, dist, order = T("knn_graph(incl self)", lambda: K.knn_graph(xy, k, True))
w32, w64, sc = T("decay_weights", lambda: K.decay_weights(dist, None, True))
xu_indptr, xu_packed = T("csr_select", lambda: K.csr_select(*csr, gm_d))
print("  nnz_u =", xu_packed.numel())
pk_rows, row_end = T("csr_select_rows (single pass)", lambda: K.csr_select_rows(*csr, gm_d))
S2, rn2 = T("diffuse_lr_score (row ranges)", lambda: K.diffuse_lr_score(csr[0], pk_rows, ncols, idx, w32, order, lig, rec, row_end=row_end))
S, row_nnz = T("diffuse_lr_score", lambda: K.diffuse_lr_score(xu_indptr, xu_packed, ncols, idx, w32, order, lig, rec))
print("  row-range result identical:", bool((S2 == S).all().item()), bool((rn2 == row_nnz).all().item()))
T("diffuse_lr_score(order=None)", lambda: K.diffuse_lr_score(xu_indptr, xu_packed, ncols, idx, w32, None, lig, rec))
Xd, neg = T("csr_select_dense", lambda: K.csr_select_dense(*csr, gm_d, ncols))
Sd, rnd = T("diffuse_lr_score_dense", lambda: K.diffuse_lr_score_dense(Xd, neg, ncols, idx, w32, order, lig, rec))
T("diffuse_lr_score_dense(order=None)", lambda: K.diffuse_lr_score_dense(Xd, neg, ncols, idx, w32, None, lig, rec))
print("  dense == sparse structure:", bool(((Sd != 0) == (S != 0)).all().item()), " max rel diff:",
      float(((Sd - S).abs() / S.abs().clamp_min(1e-30)).max().item()), " ldX =", Xd.shape[1])
T("csr_compact", lambda: K.csr_compact(S, P, row_nnz))
idx2, dist2, order2 = T("knn_graph(excl self)", lambda: K.knn_graph(xy, k, False))
w2, _, _ = K.decay_weights(dist2, 100.0)
T("spatial_stats(observed)", lambda: K.spatial_stats(S, P, idx2, w2, False, order2))
T("spatial_stats(observed,order=None)", lambda: K.spatial_stats(S, P, idx2, w2, False, None))
cols, wperm = T("null_graph_philox", lambda: K.null_graph_philox(c.n, k, w2, 0, 27))
T("spatial_stats(null)", lambda: K.spatial_stats(S, P, cols, wperm, True, None))
T("celltype_nonzero_count", lambda: K.celltype_nonzero_count(S, P, labels, c.C))
N_ = T("neighborhood_composition", lambda: K.neighborhood_composition(labels, idx2, w2, c.C))
T("celltype_pair_contract(tcgen05)", lambda: K.celltype_pair_contract(S, P, N_, 0, order=order2), reps=3)
T("celltype_pair_contract(fp32)", lambda: K.celltype_pair_contract(S, P, N_, 1), reps=2)
T("onehot_contract_csr", lambda: K.onehot_contract_csr(xu_indptr, xu_packed, len(genes), labels, c.C))
rng = np.random.default_rng(0)
table = K.to_device(rng.random((c.C * c.C, P)), torch.float64)
bg = K.to_device(np.stack([rng.permutation(P)[:30] for _ in range(P)]).astype(np.int32), torch.int32)
T("pvalue_counts(N=1000, standard)", lambda: K.pvalue_counts(table, bg, 1000, seed=27, only_ge=True))
T("pvalue_counts(N=1000, conditional)", lambda: K.pvalue_counts(table, bg, 1000, seed=27))
pv = K.to_device(np.round(rng.random((c.C * c.C, P)), 6), torch.float64); sel = K.to_device(np.ones((c.C * c.C, P), np.uint8), torch.uint8)
T("bh_fdr", lambda: K.bh_fdr(pv, sel))
