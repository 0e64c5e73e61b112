"""Where the end-to-end time of prepareLRInteraction goes (pinned host inputs, config 2): synchronised phase timers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse as sp, pandas as pd
import laris_b200 as la
from laris_b200 import engine, _kernels as K
from laris_b200.tools import _utils
from laris_b200._anndata import make_anndata
from laris_b200.synth import make_dataset
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a, lr, c = make_dataset(cfg)
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
Xh = sp.csr_matrix((pin(a.X.data.astype(np.float32)), pin(a.X.indices.astype(np.int32)), pin(a.X.indptr.astype(np.int64))), shape=a.X.shape)
a.X = Xh
a.obsm["X_spatial"] = pin(a.obsm["X_spatial"])
for _ in range(4):
    out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("whole call ms:", [round(1e3 * t, 2) for t in ts])
def T(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {name:38s} {1e3 * (time.perf_counter() - t0):7.3f} ms", flush=True)
    return r
for rep in range(2):
    print("pass", rep)
    xy = T("coords", lambda: _utils._coords(a, "X_spatial"))
    lg, rg = T("gene indices", lambda: la.tl._lr_gene_indices(a, lr))
    graph = T("build_graph (H2D xy + K1)", lambda: engine.build_graph(xy, c.k, True, None))
    csr = T("upload_csr (H2D)", lambda: engine.upload_csr(a.X))
    xu, lc, rc = T("select_lr_genes", lambda: engine.select_lr_genes(a.X, lg, rg, csr_dev=csr))
    sm = T("lr_scores (K2+K3)", lambda: engine.lr_scores(xu, graph, lc, rc))
    comp = T("csr_compact", lambda: K.csr_compact(sm.S, sm.P, sm.row_nnz))
    host = T("to_host (D2H, pinned)", lambda: K.to_host(*comp))
    X = T("scores_to_csr (all of the above 2)", lambda: engine.scores_to_csr(sm, np.float32))
    names = T("lr names", lambda: (lr["ligand"].astype(str) + "::" + lr["receptor"].astype(str)).to_numpy(dtype=object))
    T("obs.copy + obsm copy + make_anndata", lambda: make_anndata(X, obs=a.obs.copy(), var=pd.DataFrame(index=pd.Index(names, dtype=object)),
                                                                   obsm={k: np.array(v, copy=True) for k, v in a.obsm.items()}))
