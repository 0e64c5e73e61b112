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
        run(300, 40, 8); run(1000, 128, 9); run(5000, 200, 12); run(4097, 500, 20); run(20000, 131, 40, timing=True)
    elif mode == "big":
        run(100000, 500, 20, timing=True); run(1000000, 2000, 40, dens=0.2, timing=True)
    else:                                   # the synthetic configs: real score matrix, real neighbourhood composition
        from laris_b200 import engine
        from laris_b200.synth import make_dataset
        import laris_b200 as la
        for cfg in (2, 3):
            a, lr, c = make_dataset(cfg)
            lg, rg = la.tl._lr_gene_indices(a, lr)
            g1 = engine.build_graph(a.obsm["X_spatial"], c.k, True, None)
            xu, lc, rc = engine.select_lr_genes(a.X, lg, rg)
            sm = engine.lr_scores(xu, g1, lc, rc)
            g2 = engine.build_graph(a.obsm["X_spatial"], c.k, False, None)
            lab = K.to_device(a.obs["CellTypes"].cat.codes.to_numpy().astype(np.int32), torch.int32)
            N_ = K.neighborhood_composition(lab, g2.idx, g2.w, c.C)
            print(f"config {cfg}: n={c.n} P={c.P} C={c.C}  nonzero fraction of N = {float((N_ != 0).float().mean()):.3f}", flush=True)
            ref, _ = K.celltype_pair_contract(sm.S, c.P, N_, 1)
            for name, order in (("natural order", None), ("spatial order", g2.order)):
                num, _ = K.celltype_pair_contract(sm.S, c.P, N_, 0, order=order)
                r, x = ref.cpu().numpy(), num.cpu().numpy()
                m = np.abs(r) > 1e-3 * np.abs(r).max()
                ts = []
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); K.celltype_pair_contract(sm.S, c.P, N_, 0, order=order); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                print(f"   tcgen05 {name}: max rel {np.abs(x - r)[m].max() and (np.abs(x - r)[m] / np.abs(r)[m]).max():.2e}  {min(ts):.3f} ms"
                      f" -> {2.0*c.n*c.P*c.C*c.C/min(ts)/1e9:.1f} TFLOP/s algorithmic", flush=True)
