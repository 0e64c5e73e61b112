"""Host-side profile of the public prepareLRInteraction call (diagnostic)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laris_b200 as la
from laris_b200.synth import make_dataset
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a, lr, c = make_dataset(cfg)
for _ in range(3):
    out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
torch.cuda.synchronize()
t0 = time.perf_counter(); out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k); torch.cuda.synchronize()
print("e2e ms", 1e3 * (time.perf_counter() - t0), "nnz", out.X.nnz)
pr = cProfile.Profile(); pr.enable()
out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
