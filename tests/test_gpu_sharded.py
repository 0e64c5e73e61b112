"""the pair- / permutation-sharded run gives the single-GPU result.

Two ranks (gloo for the host-side gathers) share cuda:0 - the data-path kernels have no collective and never wait
on one another, so the ranks only meet in the all-gathers of per-pair vectors / tables and the all-reduce of the
exceedance counts, exactly as they do over NCCL on two GPUs (tools/dist_check.py runs the same check there)."""
import os
import pickle
import socket
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(n_nearest_neighbors=8, n_repeats=2, number_nearest_neighbors=8, n_permutations=60, n_cells_expressed_threshold=10,
          verbose=False)


def _dataset():
    from laris_b200.synth import make_dataset
    return make_dataset(1, n=3000, G=200, G_u=100, P=61, C=4)


def _run(sharded, rng):
    import laris_b200 as la
    a, lr, c = _dataset()
    if sharded:
        lr_adata = la.dist.prepare_sharded(a, lr, number_nearest_neighbors=8)
    else:
        lr_adata = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8)
    laris_lr, table = la.tl.runLARIS(lr_adata, a, n_top_lr=len(lr), rng=rng, **KW)
    return dict(lr=laris_lr, table=table, var=lr_adata.var.copy(), names=list(lr_adata.var_names), nnz=int(lr_adata.X.nnz))


def _worker(rank, world, port, rng, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = _run(True, rng)
        with open(os.path.join(outdir, f"rank{rank}.pkl"), "wb") as f:
            pickle.dump(res, f)
    finally:
        dist.destroy_process_group()


def _key(df):
    return df.sort_values(["sender", "receiver", "interaction_name"]).reset_index(drop=True)


@pytest.mark.parametrize("rng", ["philox", "numpy"])
def test_two_rank_run_equals_single(rng):
    import torch.multiprocessing as mp
    single = _run(False, rng)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, rng, d), nprocs=2, join=True)
        ranks = [pickle.load(open(os.path.join(d, f"rank{r}.pkl"), "rb")) for r in range(2)]
    # the pair blocks partition the score matrix
    assert ranks[0]["names"] + ranks[1]["names"] == single["names"]
    assert ranks[0]["nnz"] + ranks[1]["nnz"] == single["nnz"]
    var = np.concatenate([r["var"]["LR_SpatialSpecificity"].to_numpy() for r in ranks])
    assert np.allclose(var, single["var"]["LR_SpatialSpecificity"].to_numpy(), rtol=1e-12, atol=1e-15)
    for r in ranks:                                            # every rank ends with the global answer
        assert list(r["lr"].index) == list(single["lr"].index)
        assert np.allclose(r["lr"]["score"].to_numpy(), single["lr"]["score"].to_numpy(), rtol=1e-12, atol=1e-15)
        a, b = _key(r["table"]), _key(single["table"])
        assert len(a) == len(b) and (a["interaction_name"] == b["interaction_name"]).all()
        assert np.allclose(a["interaction_score"].to_numpy(), b["interaction_score"].to_numpy(), rtol=1e-9, atol=1e-14)
        assert np.array_equal(a["p_value"].to_numpy(), b["p_value"].to_numpy())      # integer counts: exact
        assert np.allclose(a["p_value_fdr"].to_numpy(), b["p_value_fdr"].to_numpy(), rtol=1e-12)
