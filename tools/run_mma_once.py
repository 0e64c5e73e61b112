"""One tcgen05 pair contraction at a fixed size (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laris_b200 import _kernels as K
n, P, C = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (200000, 1000, 30)
g = torch.Generator(device="cuda").manual_seed(0)
S = torch.rand((n, K.padded_ld(P)), device="cuda", generator=g) * (torch.rand((n, K.padded_ld(P)), device="cuda", generator=g) < 0.3)
N = torch.zeros((n, C), device="cuda")
N.scatter_(1, torch.randint(0, C, (n, 3), device="cuda", generator=g), torch.rand((n, 3), device="cuda", generator=g))
for _ in range(2):
    num, q = K.celltype_pair_contract(S.contiguous(), P, N, 0)
torch.cuda.synchronize()
print("ok", float(num.sum()))
