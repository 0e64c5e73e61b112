"""Device time and equality of the packed-row statistics pass (K4 sparse) against the dense passes (K4a tiles + K4b):
    python tools/time_k4s.py [n] [P] [k] [R] [density]"""
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
density = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
dev = torch.device("cuda", 0)
xy = torch.rand((n, 2), dtype=torch.float64, device=dev) * 10.0 * np.sqrt(n)
g = engine.build_graph(xy, k, False, 100.0)
ld = K.padded_ld(P)
S = torch.rand((n, ld), dtype=torch.float32, device=dev)
S *= (torch.rand((n, ld), device=dev) < density)
S[:, P:] = 0
S[n // 2, :P] = 1.5                                   # one dense row: the long-row path
cols, wts, _ = engine.null_graphs(g, R, 27, "philox")
sm = engine.ScoreMatrix(S, P)


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


engine.queue_row_pointers(sm)
nnz = int(sm.cache["nnz_get"]()[0])
sp_ptr = sm.cache["sp_ptr"]
t_pack, entries = timed(lambda: K.scores_pack(S, P, sp_ptr, nnz))
# the packed rows hold exactly the non-zeros of S
e = entries.cpu().numpy()
colv = (e & 0xffffffff).astype(np.int64)
val = (e >> 32).astype(np.uint32).view(np.float32) if False else ((e >> 32) & 0xffffffff).astype(np.uint32).view(np.float32)
ptr = sp_ptr.cpu().numpy()
rows = np.repeat(np.arange(n), np.diff(ptr))
chk = np.random.default_rng(0).integers(0, len(e), 200000)
chk = chk[colv[chk] < ld]                              # padding entries carry column ld + j and value 0
Sh = S[torch.from_numpy(rows[chk]).to(dev), torch.from_numpy(colv[chk]).to(dev)].cpu().numpy()
pack_ok = (bool(np.array_equal(Sh, val[chk])) and int((colv < ld).sum()) == int((S != 0).sum().item())
           and bool((val[colv >= ld] == 0).all()) and bool((np.diff(ptr) % 32 == 0).all()))
t_sp, a = timed(lambda: K.spatial_stats_sparse(S, P, sp_ptr, entries, g.idx, g.w, g.order, cols, wts))
t_obs, b0 = timed(lambda: K.spatial_stats_batched(S, P, g.idx, g.w, g.order))
t_null, b1 = timed(lambda: K.spatial_stats_batched(S, P, None, None, None, cols, wts))
ref = torch.cat([b0, b1])
err = float(((a - ref).abs() / ref.abs().clamp_min(1e-300)).max().item())
print(f"n={n} P={P} k={k} R={R} density={density}: pack {t_pack:.3f} ms (ok={pack_ok}, nnz={nnz}), sparse stats {t_sp:.3f} ms, "
      f"dense K4a {t_obs:.3f} + K4b {t_null:.3f} ms; max rel diff {err:.2e}, counts equal={bool(torch.equal(a[:, 4], ref[:, 4]))}")
