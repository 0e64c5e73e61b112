"""Host vs device timeline of one device-resident whole-path step: for every C-ABI stage the host time at which it was
enqueued and the device interval in which it ran (both in ms from the start of the step).  Shows where the device idles
waiting for the host (and vice versa).      python tools/step_timeline.py [config] [n]"""
import os
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
a, lr, c = make_dataset(cfg, n=n, fast=os.environ.get("LARIS_FAST_DATA", "0") == "1")
csr_d = engine.upload_csr(a.X)
a_dev = AnnDataLite(a.X, a.obs, a.var, {"X_spatial": K.to_device(a.obsm["X_spatial"], torch.float64)})
kw = dict(n_nearest_neighbors=c.k, n_repeats=c.R, n_top_lr=c.P, number_nearest_neighbors=c.k, n_permutations=c.N,
          verbose=False, rng="philox", output="arrays")


def step():
    ad = la.tl.prepareLRInteraction(a_dev, lr, number_nearest_neighbors=c.k, csr_dev=csr_d, to_host=False)
    return la.tl.runLARIS(ad, a_dev, **kw)


for _ in range(3):
    step()
torch.cuda.synchronize()


def one():
    with K.timing() as t:
        base = torch.cuda.Event(enable_timing=True)
        h_base = time.perf_counter()
        base.record()
        step()
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        torch.cuda.synchronize()
        h_end = time.perf_counter()
    return t, base, end, h_base, h_end


def show(t, base, end, h_base, h_end):
    print(f"step: host {1e3 * (h_end - h_base):.2f} ms, device {base.elapsed_time(end):.2f} ms")
    print(f"{'stage':34s} {'host enq':>9s} {'host ret':>9s} | {'dev start':>9s} {'dev end':>9s} {'dev ms':>7s} {'idle before':>11s}")
    prev_end = 0.0
    for stage, e0, e1, h0, h1 in t.timeline:
        d0, d1 = base.elapsed_time(e0), base.elapsed_time(e1)
        print(f"{stage:34s} {1e3 * (h0 - h_base):9.2f} {1e3 * (h1 - h_base):9.2f} | {d0:9.2f} {d1:9.2f} {d1 - d0:7.2f} {d0 - prev_end:11.2f}")
        prev_end = d1


# LARIS_TIMELINE_TRIES > 1: also show the slowest of that many steps (intermittent host stalls)
tries = int(os.environ.get("LARIS_TIMELINE_TRIES", "1"))
runs = [one() for _ in range(tries)]
times = [r[1].elapsed_time(r[2]) for r in runs]
show(*runs[int(np.argsort(times)[len(times) // 2])])
if tries > 1:
    print("\nall steps (ms):", [round(x, 2) for x in times], "\nslowest:")
    show(*runs[int(np.argmax(times))])
