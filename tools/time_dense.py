"""Time the dense-panel K2+K3 kernel (and check it against the sparse kernel).  usage: time_dense.py [cfg] [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laris_b200 import _kernels as K, engine
from laris_b200.synth import make_dataset
import laris_b200 as la
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
a, lr, c = make_dataset(cfg, n=n)
xy = K.to_device(a.obsm["X_spatial"], torch.float64)
csr = engine.upload_csr(a.X)
lg, rg = la.tl._lr_gene_indices(a, lr)
plan = engine.plan_lr_table(lg, rg, a.X.shape[1])
g = engine.build_graph(xy, c.k, True, None)
xd = engine.select_planned(plan, csr, c.n, "dense")
xs = engine.select_planned(plan, csr, c.n, "sparse")
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.fill_(1.0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return np.median(ts), min(ts), r
md, mn, smd = t(lambda: engine.lr_scores(xd, g, plan.lig_d, plan.rec_d))
ms, mns, sms = t(lambda: engine.lr_scores(xs, g, plan.lig_d, plan.rec_d))
G_u = len(plan.genes)
alg = 4 * c.n * G_u + 8 * c.n * c.k + 4 * c.n * c.P
same = bool(((smd.S != 0) == (sms.S != 0)).all().item())
rel = float(((smd.S - sms.S).abs() / sms.S.abs().clamp_min(1e-30)).max().item())
print(f"cfg{cfg} n={c.n}: dense med {md*1e3:.1f} us min {mn*1e3:.1f} us -> {alg/md/1e6:.0f} GB/s ({alg/md/1e6/6455.6:.3f} of peak) | sparse med {ms*1e3:.1f} us"
      f" | structure equal {same} max rel diff {rel:.2e} row_nnz equal {bool((smd.row_nnz == sms.row_nnz).all().item())}", flush=True)
