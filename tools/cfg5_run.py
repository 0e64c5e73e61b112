"""Config 5 (Stereo-seq-like bin lattice): radius graph, LR scores on it, cell-type neighbourhood aggregation.
usage: cfg5_run.py [n]   - scale check of the radius path (north-star config 5), timings printed per stage."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laris_b200 as la
from laris_b200 import engine, _kernels as K
from laris_b200.synth import make_dataset
n = int(sys.argv[1]) if len(sys.argv) > 1 else None
t0 = time.perf_counter()
a, lr, c = make_dataset(5, n=n)
print(f"config 5: n={c.n} G={c.G} P={c.P} C={c.C} (synthesised in {time.perf_counter()-t0:.1f} s)", flush=True)
radius = 1.5                                   # lattice pitch 1: the 8-neighbourhood plus the bin itself
def T(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {name:46s} {1e3*(time.perf_counter()-t):10.2f} ms", flush=True)
    return r
for rep in range(2):
    print("pass", rep, flush=True)
    xy = K.to_device(a.obsm["X_spatial"], torch.float64)
    g = T("radius graph (count+scan+fill+weights+pad)", lambda: engine.build_radius_graph(xy, radius, True))
    deg = (g.csr[0][1:] - g.csr[0][:-1])
    print(f"    nnz={int(g.csr[0][-1])} mean degree={float(deg.float().mean()):.2f} max={int(deg.max())}", flush=True)
    lab = K.to_device(a.obs["CellTypes"].cat.codes.to_numpy().astype(np.int32), torch.int32)
    N_ = T("neighbourhood aggregation (n x C)", lambda: K.neighborhood_composition(lab, g.idx, g.w, c.C))
    out = T("prepareLRInteraction(radius=1.5) host to host", lambda: la.tl.prepareLRInteraction(a, lr, radius=radius))
    print(f"    scores nnz={out.X.nnz}", flush=True)
    del out, g, N_
    torch.cuda.empty_cache()
