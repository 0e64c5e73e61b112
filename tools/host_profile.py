"""Where the host time of one whole-path step goes (cProfile of the device-resident step and of the public-API step).

    python tools/host_profile.py [config] [n]
"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import laris_b200 as la
from laris_b200 import _kernels as K, engine
from laris_b200._anndata import AnnDataLite
from laris_b200.synth import make_dataset

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
a, lr, c = make_dataset(cfg, n=n)
kw = dict(n_nearest_neighbors=c.k, n_repeats=c.R, n_top_lr=c.P, number_nearest_neighbors=c.k, n_permutations=c.N,
          verbose=False, rng="philox")
csr_d = engine.upload_csr(a.X)
a_dev = AnnDataLite(a.X, a.obs, a.var, {"X_spatial": K.to_device(a.obsm["X_spatial"], torch.float64)})


def device_step():
    ad = la.tl.prepareLRInteraction(a_dev, lr, number_nearest_neighbors=c.k, csr_dev=csr_d, to_host=False)
    out = la.tl.runLARIS(ad, a_dev, output="arrays", **kw)
    torch.cuda.synchronize()
    return out


def host_step():
    ad = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    out = la.tl.runLARIS(ad, a, **kw)
    torch.cuda.synchronize()
    return out


for name, fn in (("device-resident step", device_step), ("public API step", host_step)):
    for _ in range(2):
        fn()
    t0 = time.perf_counter()
    fn()
    print(f"## {name}: {1e3 * (time.perf_counter() - t0):.1f} ms")
    pr = cProfile.Profile()
    pr.enable()
    fn()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
    print(s.getvalue())
