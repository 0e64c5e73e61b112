"""Two device-resident whole-path steps of a config (the step bench.py times), for ncu captures:
    python tools/run_step_once.py [config] [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import laris_b200 as la
from laris_b200 import _kernels as K, engine
from laris_b200._anndata import AnnDataLite
from laris_b200.synth import make_dataset

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
a, lr, c = make_dataset(cfg, n=n, fast=os.environ.get("LARIS_FAST_DATA", "0") == "1")
csr_d = engine.upload_csr(a.X)
a_dev = AnnDataLite(a.X, a.obs, a.var, {"X_spatial": K.to_device(a.obsm["X_spatial"], torch.float64)})
kw = dict(n_nearest_neighbors=c.k, n_repeats=c.R, n_top_lr=c.P, number_nearest_neighbors=c.k, n_permutations=c.N,
          verbose=False, rng="philox", output="arrays")
for _ in range(2):
    ad = la.tl.prepareLRInteraction(a_dev, lr, number_nearest_neighbors=c.k, csr_dev=csr_d, to_host=False)
    la.tl.runLARIS(ad, a_dev, **kw)
    torch.cuda.synchronize()
print("ok")
