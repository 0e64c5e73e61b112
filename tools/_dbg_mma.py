import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laris_b200 import _kernels as K
n, P, C = 9000, 131, 40
rng = np.random.default_rng(n + P + C)
ld = K.padded_ld(P)
S = np.zeros((n, ld), np.float32)
S[:, :P] = (rng.random((n, P)) < 0.3) * rng.gamma(2.0, 1.0, (n, P)).astype(np.float32)
N = np.zeros((n, C), np.float32)
for _ in range(3):
    N[np.arange(n), rng.integers(0, C, n)] += rng.random(n).astype(np.float32)
N[::50] = 0.0
Sd, Nd = torch.from_numpy(S).cuda(), torch.from_numpy(N).cuda()
num0, q0 = K.celltype_pair_contract(Sd, P, Nd, 0)
Q = (N[:, :, None].astype(np.float64) * N[:, None, :]).reshape(n, C * C)
ref = S[:, :P].astype(np.float64).T @ Q
a0 = num0.cpu().numpy()
nz = ref != 0
rel = np.zeros_like(ref); rel[nz] = (a0[nz] - ref[nz]) / ref[nz]
idx = np.argsort(-np.abs(rel).ravel())[:8]
for i in idx:
    p, ab = divmod(i, C * C)
    terms = np.nonzero(S[:, p] * Q[:, ab])[0]
    print(f"p={p} ab={ab} ref={ref[p,ab]:.6e} got={a0[p,ab]:.6e} rel={rel[p,ab]:.3e} nterms={len(terms)}")
    if len(terms) <= 3:
        for t in terms: print("    cell", t, "S", S[t, p], "Na", N[t, ab // C], "Nb", N[t, ab % C], "Q", Q[t, ab])
print("count > 1e-5:", int((np.abs(rel) > 1e-5).sum()), "of", int(nz.sum()))
