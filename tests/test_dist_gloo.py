"""Host-side sharding logic of the multi-GPU path on CPU: world_size 2, gloo backend.

The data-path kernels are single-GPU and have no collective; what the N > 1 path adds is the pair /
permutation partition and the gathers of per-pair vectors, tables and exceedance counts
(laris_b200/dist.py).  Those run here over gloo exactly as they run over NCCL on the GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, P, C):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from laris_b200 import dist as ld
        names = np.array([f"L{p}::R{p}" for p in range(P)], dtype=object)
        sh = ld.Shard(rank, world, P, names)
        b = ld.pair_bounds(P, world)
        assert sh.start == b[rank] and sh.stop == b[rank + 1]
        # per-pair vectors: every rank contributes its block, everyone ends with the global vector
        full = np.arange(P, dtype=np.float64) * 1.5 + 0.25
        got = ld.all_gather_pairs(full[sh.start:sh.stop], sh)
        assert got.shape == (P,) and np.array_equal(got, full)
        # (groups x pairs) tables gather along the pair axis
        table = np.arange(C * C * P, dtype=np.float64).reshape(C * C, P)
        got = ld.all_gather_pairs(table[:, sh.start:sh.stop], sh, axis=1)
        assert np.array_equal(got, table)
        # [5,P] statistics block (leading axis kept)
        stats = np.random.default_rng(3).random((5, P))
        assert np.array_equal(ld.all_gather_pairs(stats[:, sh.start:sh.stop], sh, axis=-1), stats)
        # exceedance counts of a permutation shard add up: counts are integers, the sum is exact
        pb = ld.perm_bounds(1000, world)
        rng = np.random.default_rng(11)
        draws = rng.integers(0, 2, size=(1000, C * C, 7))                 # same stream on every rank
        mine = draws[pb[rank]:pb[rank + 1]].sum(0).astype(np.int64)
        total = ld.all_reduce_sum(mine)
        assert np.array_equal(total, draws.sum(0))
        # the tensor versions the device-resident pipeline uses (CUDA tensors over NCCL; host tensors over gloo here)
        st = torch.from_numpy(np.random.default_rng(4).random((3, 5, P)))
        got = ld.all_gather_pairs_dev(st[:, :, sh.start:sh.stop].contiguous(), sh, axis=-1)
        assert got.shape == st.shape and torch.equal(got, st)
        rows = torch.arange(P * C, dtype=torch.int64).reshape(P, C)
        assert torch.equal(ld.all_gather_pairs_dev(rows[sh.start:sh.stop].contiguous(), sh, axis=0), rows)
        cnt = torch.from_numpy(draws[pb[rank]:pb[rank + 1]].sum(0).astype(np.int32))
        assert np.array_equal(ld.all_reduce_sum_dev(cnt).numpy(), draws.sum(0))
        # row partition of the expression matrix by stored entries
        indptr = np.concatenate([[0], np.cumsum(np.random.default_rng(1).integers(0, 9, 1000))])
        rb = ld.row_bounds_by_nnz(indptr, 4)
        assert rb[0] == 0 and rb[-1] == 1000 and (np.diff(rb) >= 0).all()
        assert np.abs(np.diff(indptr[rb]) - indptr[-1] / 4).max() <= 9
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,C", [(11, 3), (500, 4)])
def test_pair_and_permutation_sharding_world2(P, C):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, P, C), nprocs=2, join=True)


def test_bounds_cover_and_balance():
    from laris_b200.dist import pair_bounds, perm_bounds
    for P in (1, 2, 7, 500, 2001):
        for w in (1, 2, 4, 8):
            b = pair_bounds(P, w)
            sizes = np.diff(b)
            assert b[0] == 0 and b[-1] == P and (sizes >= 0).all() and sizes.max() - sizes.min() <= 1
    assert np.array_equal(perm_bounds(1000, 8), np.arange(0, 1001, 125))
