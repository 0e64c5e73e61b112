"""Device time of the batched null pass (K4b): [LARIS_B200_K4B=hbm] python tools/time_k4b.py [n] [P] [k] [R]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from laris_b200 import _kernels as K, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
R = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
xy = torch.rand((n, 2), dtype=torch.float64, device=dev) * 10.0 * np.sqrt(n)
g = engine.build_graph(xy, k, False, 100.0)
S = torch.rand((n, K.padded_ld(P)), dtype=torch.float32, device=dev)
S *= (torch.rand((n, K.padded_ld(P)), device=dev) < 0.25)
cols, wts, _ = engine.null_graphs(g, R, 27, "philox")


def timed(fn, reps=4):
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


t_new, a = timed(lambda: K.spatial_stats_batched(S, P, None, None, None, cols, wts))
ref = torch.stack([K.spatial_stats(S, P, cols[r].contiguous(), wts[r].contiguous(), True, None) for r in range(R)])
ok = bool(torch.allclose(a, ref, rtol=2e-6, atol=1e-300))
gb = (4.0 * (1 + R * k) * n * P + 8.0 * n * k * R) / 1e9
print(f"n={n} P={P} k={k} R={R} mode={os.environ.get('LARIS_B200_K4B', 'l2')}: {t_new:.3f} ms "
      f"({gb / t_new * 1e3:.0f} GB/s of algorithmic bytes), equal to the per-repeat kernel={ok}")
