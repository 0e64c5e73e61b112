"""Time laris_diffuse_lr_score under tile-shape overrides (diagnostic; LARIS_DIFFUSE_* env)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laris_b200 import _kernels as K, engine
from laris_b200.synth import make_dataset
import laris_b200 as la

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
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
idx, dist, order = K.knn_graph(xy, k, True)
w32, _, _ = K.decay_weights(dist, None, False)
xu_indptr, xu_packed = K.csr_select(*csr, gm_d)
nnz_u = xu_packed.numel()
alg = 8 * nnz_u + 8 * c.n * k + 4 * c.n * P + 8 * (c.n + 1)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
print(f"config {cfg}: n={c.n} G_u={len(genes)} P={P} nnz_u={nnz_u} alg_bytes={alg/1e6:.1f} MB")
ref = None
combos = [(0, 0, 0)] + [(tc, pch, nt) for tc in (32, 16) for pch in (128, 64, 32) for nt in (256, 512, 1024)]
for tc, pch, nt in combos:
    for kk, v in (("LARIS_DIFFUSE_TC", tc), ("LARIS_DIFFUSE_PCH", pch), ("LARIS_DIFFUSE_NT", nt)):
        if v: os.environ[kk] = str(v)
        else: os.environ.pop(kk, None)
    try:
        ts = []
        for it in range(6):
            flush.fill_(1.0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S, rn = K.diffuse_lr_score(xu_indptr, xu_packed, ncols, idx, w32, order, lig, rec); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        if ref is None: ref = (S.clone(), rn.clone())
        ok = torch.equal(S, ref[0]) and torch.equal(rn, ref[1])
        t = min(ts[1:])
        print(f"TC={tc:2d} PCH={pch:3d} NT={nt:4d}: min {t:8.3f} ms  med {np.median(ts[1:]):8.3f} ms  {alg/t/1e6:7.1f} GB/s  same={ok}", flush=True)
    except Exception as ex:
        print(f"TC={tc} PCH={pch} NT={nt}: {str(ex)[:100]}", flush=True)
