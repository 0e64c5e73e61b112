"""tcgen05 pair contraction (impl=0) vs the fp32 CUDA-core kernel (impl=1) and a float64 numpy contraction."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laris_b200 import _kernels as K

def run(n, P, C, dens=0.3, seed=0, timing=False):
    rng = np.random.default_rng(seed)
    ld = K.padded_ld(P)
    S = np.zeros((n, ld), np.float32)
    S[:, :P] = (rng.random((n, P)) < dens) * rng.gamma(2.0, 1.0, (n, P)).astype(np.float32)
    N = np.zeros((n, C), np.float32)
    for _ in range(3):
        N[np.arange(n), rng.integers(0, C, n)] += rng.random(n).astype(np.float32)
    Sd, Nd = torch.from_numpy(S).cuda(), torch.from_numpy(N).cuda()
    num0, q0 = K.celltype_pair_contract(Sd, P, Nd, 0)
    torch.cuda.synchronize()
    num1, q1 = K.celltype_pair_contract(Sd, P, Nd, 1)
    torch.cuda.synchronize()
    if n * P * C * C <= 4e9:
        Q = (N[:, :, None].astype(np.float64) * N[:, None, :]).reshape(n, C * C)
        ref = S[:, :P].astype(np.float64).T @ Q
    else:
        ref = num1.cpu().numpy()
    a0, a1 = num0.cpu().numpy(), num1.cpu().numpy()
    scale = np.abs(ref).max()
    nzm = np.abs(ref) > 1e-3 * scale
    e0 = np.abs(a0 - ref)[nzm] / np.abs(ref)[nzm]
    e1 = np.abs(a1 - ref)[nzm] / np.abs(ref)[nzm]
    print(f"n={n} P={P} C={C}: tcgen05 max rel {e0.max():.2e} (abs/scale {np.abs(a0-ref).max()/scale:.2e}) | fp32 kernel max rel {e1.max():.2e}"
          f" | zero pattern equal {np.array_equal(a0 == 0, ref == 0)}", flush=True)
    if timing:
        for impl in (0, 1):
            ts = []
            for _ in range(3):
                e_0, e_1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e_0.record(); K.celltype_pair_contract(Sd, P, Nd, impl); e_1.record(); torch.cuda.synchronize()
                ts.append(e_0.elapsed_time(e_1))
            print(f"   impl={impl}: {min(ts):.3f} ms  -> {2.0*n*P*C*C/min(ts)/1e9:.1f} TFLOP/s algorithmic", flush=True)
    return e0.max()

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "small"
    if mode == "small":
        run(300, 40, 5); run(1000, 128, 8); run(5000, 200, 12); run(4097, 500, 20); run(20000, 131, 40, timing=True)
    else:
        run(100000, 500, 20, timing=True); run(1000000, 2000, 40, dens=0.2, timing=True)
