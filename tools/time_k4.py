"""Device time of the observed-statistics pass (tiled K4a, both staging modes) and of the legacy kernel on synthetic
scores of a given shape:  [LARIS_B200_K4A_STAGING=bulk|gather4|cpasync] python tools/time_k4.py [n] [P] [k]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from laris_b200 import _kernels as K, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda", 0)
xy = torch.rand((n, 2), dtype=torch.float64, device=dev) * 10.0 * np.sqrt(n)
g = engine.build_graph(xy, k, False, 100.0)
S = torch.rand((n, K.padded_ld(P)), dtype=torch.float32, device=dev)
S *= (torch.rand((n, K.padded_ld(P)), device=dev) < 0.25)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), out


mode = os.environ.get("LARIS_B200_K4A_STAGING", "cpasync")
t_new, a = timed(lambda: K.spatial_stats_batched(S, P, g.idx, g.w, g.order))
t_old, b = timed(lambda: K.spatial_stats(S, P, g.idx, g.w, False, g.order))
ok = bool(torch.allclose(a[0], b, rtol=1e-5))
gb = (4 * n * P + 8 * n * k) / 1e9
print(f"n={n} P={P} k={k} staging={mode}: tiled {t_new:.3f} ms ({gb / t_new * 1e3:.0f} GB/s algorithmic), "
      f"legacy {t_old:.3f} ms, equal={ok}")
