"""Run the M1 step stages once or twice on a config (profiling target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laris_b200 import _kernels as K, engine
from laris_b200.synth import make_dataset
import laris_b200 as la
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
a, lr, c = make_dataset(cfg)
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
dense = os.environ.get("LARIS_DENSE", "1") == "1"
for _ in range(reps):
    idx, dist, order = K.knn_graph(xy, k, True)
    w32, _, _ = K.decay_weights(dist, None, False)
    if dense:
        Xd, neg = K.csr_select_dense(*csr, gm_d, ncols)
        S, rn = K.diffuse_lr_score_dense(Xd, neg, ncols, idx, w32, order, lig, rec)
    else:
        xu_indptr, xu_packed = K.csr_select(*csr, gm_d)
        S, rn = K.diffuse_lr_score(xu_indptr, xu_packed, ncols, idx, w32, order, lig, rec)
torch.cuda.synchronize()
print("ok", float(S.sum()))
