"""Multi-GPU sharding of the LR scoring path (one process per GPU, torch.distributed).

The path shards with no halo: coordinates and the kNN graph are
replicated, **LR pairs** (columns of S) partition every per-pair stage, and the
**permutation axis** partitions the p-value null (the device RNG is keyed by
global ids, so the counts are independent of the split).  The only exchanges
are all-gathers of per-pair vectors / (P x C^2) tables and one all-reduce of the
exceedance counts - a few MB over NVLink, never on the critical path.

Usage (torchrun, backend nccl)::

    lr_local = la.dist.prepare_sharded(adata, lr_df)          # this rank's pairs
    laris_lr, table = la.tl.runLARIS(lr_local, adata, ...)    # identical result on every rank

``runLARIS`` recognises the shard record in ``lr_local.uns['laris_b200_shard']``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SHARD_KEY = "laris_b200_shard"


@dataclass
class Shard:
    rank: int
    world: int
    P_global: int
    names_global: np.ndarray        # all P "ligand::receptor" names, global order

    @property
    def bounds(self):
        return pair_bounds(self.P_global, self.world)

    @property
    def start(self):
        return int(self.bounds[self.rank])

    @property
    def stop(self):
        return int(self.bounds[self.rank + 1])


def pair_bounds(P: int, world: int) -> np.ndarray:
    """Contiguous, balanced partition of [0,P): rank r owns [b[r], b[r+1])."""
    base, extra = divmod(P, world)
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:extra] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


def perm_bounds(n_perm: int, world: int) -> np.ndarray:
    """Partition of the permutation axis [0, n_perm) (same rule as pairs)."""
    return pair_bounds(n_perm, world)


def _dist():
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun / init_process_group)")
    return dist


def _carrier(arr):
    """numpy -> tensor on the device the process group communicates over."""
    import torch
    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    return t


def all_gather_pairs(local: np.ndarray, shard: Shard, axis: int = -1) -> np.ndarray:
    """Concatenate every rank's pair block along ``axis`` in global pair order."""
    import torch
    dist = _dist()
    b = shard.bounds
    width = int((b[1:] - b[:-1]).max())
    x = np.moveaxis(np.asarray(local), axis, 0)
    pad = np.zeros((width,) + x.shape[1:], dtype=x.dtype)
    pad[: x.shape[0]] = x
    mine = _carrier(pad)
    outs = [torch.empty_like(mine) for _ in range(shard.world)]
    dist.all_gather(outs, mine)
    parts = [o.cpu().numpy()[: int(b[r + 1] - b[r])] for r, o in enumerate(outs)]
    return np.moveaxis(np.concatenate(parts, axis=0), 0, axis)


def all_reduce_sum(arr: np.ndarray) -> np.ndarray:
    dist = _dist()
    t = _carrier(arr)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def prepare_sharded(adata, lr_df, number_nearest_neighbors: int = 10, use_rep_spatial: str = "X_spatial", **kw):
    """``prepareLRInteraction`` over this rank's block of LR pairs.

    Every rank builds the (replicated) graph and scores only its own pairs; the
    returned AnnData holds the n x P_local block plus the shard record."""
    from .tools import prepareLRInteraction
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    P = len(lr_df)
    names = (lr_df["ligand"].astype(str) + "::" + lr_df["receptor"].astype(str)).to_numpy(dtype=object)
    sh = Shard(rank, world, P, names)
    local = prepareLRInteraction(adata, lr_df.iloc[sh.start:sh.stop].reset_index(drop=True),
                                 number_nearest_neighbors, use_rep_spatial, **kw)
    local.uns[SHARD_KEY] = sh
    return local


def shard_of(lr_adata):
    sh = lr_adata.uns.get(SHARD_KEY) if hasattr(lr_adata, "uns") else None
    return sh if isinstance(sh, Shard) else None
