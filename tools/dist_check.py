"""NCCL check of the sharded path: torchrun --nproc-per-node N tools/dist_check.py [cfg] [n]

Every rank scores its block of LR pairs on its own GPU (laris_b200.dist.prepare_sharded), runLARIS gathers per-pair
vectors / tables over NCCL and splits the permutations; rank 0 also runs the single-GPU path and compares."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import laris_b200 as la
from laris_b200.synth import make_dataset

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
a, lr, c = make_dataset(cfg, n=n)
kw = dict(n_nearest_neighbors=c.k, n_repeats=c.R, n_top_lr=c.P, number_nearest_neighbors=c.k, n_permutations=max(c.N, 100),
          rng="philox", verbose=False)
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    lr_local = la.dist.prepare_sharded(a, lr, number_nearest_neighbors=c.k)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    laris_lr, table = la.tl.runLARIS(lr_local, a, **kw)
    torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
    if rank == 0:
        print(json.dumps({"world": world, "pass": rep, "prepare_s": t1 - t0, "runLARIS_s": t2 - t1, "total_s": t2 - t0,
                          "pairs_local": int(lr_local.shape[1])}), flush=True)
if rank == 0:
    full = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    ref_lr, ref_table = la.tl.runLARIS(full, a, **kw)
    key = ["sender", "receiver", "interaction_name"]
    x, y = table.sort_values(key).reset_index(drop=True), ref_table.sort_values(key).reset_index(drop=True)
    ok = (list(laris_lr.index) == list(ref_lr.index)
          and np.allclose(laris_lr["score"], ref_lr["score"], rtol=1e-12, atol=1e-15)
          and np.allclose(x["interaction_score"], y["interaction_score"], rtol=1e-9, atol=1e-14)
          and np.array_equal(x["p_value"].to_numpy(), y["p_value"].to_numpy()))
    print("sharded == single-GPU:", ok, flush=True)
    assert ok
dist.destroy_process_group()
