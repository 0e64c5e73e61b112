"""Multi-GPU sharding of the LR scoring path (one process per GPU, torch.distributed).

The path shards with no halo: coordinates and the kNN graph are
replicated, **LR pairs** (columns of S) partition every S-sized pass (K2+K3, K4a,
K4b, K5, K5b), and the **permutation axis** partitions the p-value null (the device
RNG is keyed by global ids, so the counts are independent of the split).  The
exchanges are all-gathers of per-pair statistics ([5+R, P] fp64), of the K5b
counts ([C, P] int64) and of the P x C^2 co-localisation table, and one all-reduce
of the integer exceedance counts - device tensors over NCCL / NVLink, a few MB to
a few tens of MB; the P x C^2-sized table arithmetic after them is replicated.

The expression matrix is replicated through NVLink rather than PCIe:
``upload_csr_replicated`` lets every rank copy 1/world of the rows from its host
and broadcasts the blocks, so a rank's host->device traffic is nnz/world.

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


def _on_wire(t):
    """Tensor as the process group can move it: CUDA tensors go as they are over NCCL, through the host over gloo."""
    return t if _dist().get_backend() == "nccl" else t.cpu()


def all_gather_pairs_dev(local, shard: Shard, axis: int = -1):
    """Device version of ``all_gather_pairs``: every rank's pair block of a torch tensor, concatenated along ``axis`` in
    global pair order; the result lives on ``local``'s device.  Blocks are padded to the widest one on the wire."""
    import torch
    dist = _dist()
    b = shard.bounds
    width = int((b[1:] - b[:-1]).max())
    x = local.movedim(axis, 0).contiguous()
    pad = torch.zeros((width,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    mine = _on_wire(pad)
    out = torch.empty((shard.world * width,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    out = out.to(local.device).view((shard.world, width) + tuple(mine.shape[1:]))
    parts = [out[r, : int(b[r + 1] - b[r])] for r in range(shard.world)]
    return torch.cat(parts, dim=0).movedim(0, axis).contiguous()


def all_reduce_sum_dev(t):
    """In-place sum over ranks of a device tensor (exact for the integer count tables)."""
    dist = _dist()
    w = _on_wire(t)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    if w is not t:
        t.copy_(w)
    return t


def all_gather_pairs(local: np.ndarray, shard: Shard, axis: int = -1) -> np.ndarray:
    """Concatenate every rank's pair block along ``axis`` in global pair order (host arrays)."""
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


def row_bounds_by_nnz(indptr: np.ndarray, world: int) -> np.ndarray:
    """Row partition of a CSR matrix with (nearly) equal stored entries per block."""
    n = len(indptr) - 1
    targets = (np.arange(1, world, dtype=np.float64) * float(indptr[-1]) / world)
    cuts = np.searchsorted(indptr, targets, side="left").astype(np.int64)
    return np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)


def upload_csr_replicated(X):
    """The expression matrix on every rank's GPU with 1/world of it crossing each rank's PCIe link: rank r copies
    its block of rows (balanced by stored entries) from its own host copy, then the blocks are broadcast in place
    over NCCL (NVLink / NVSwitch).  Every rank must hold the same host matrix (they all read the same AnnData).
    Falls back to a plain full upload for a single rank or a non-NCCL backend."""
    import scipy.sparse as sp
    import torch
    from . import _kernels as K
    from . import engine
    dist = _dist()
    world = dist.get_world_size()
    if world == 1 or dist.get_backend() != "nccl":
        return engine.upload_csr(X)
    if not sp.issparse(X):
        raise TypeError("adata.X must be a scipy sparse matrix (as in the reference, laris/tools/__init__.py:100)")
    X = X.tocsr()
    if not X.has_canonical_format:
        X = X.copy()
        X.sum_duplicates()
    rank = dist.get_rank()
    dev = K.require_cuda()
    rb = row_bounds_by_nnz(X.indptr, world)
    eb = np.asarray(X.indptr)[rb].astype(np.int64)                       # entry offsets of the row blocks
    nnz = int(eb[-1])
    indptr = K.to_device(X.indptr, torch.int64)
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    data = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
    s, e = int(eb[rank]), int(eb[rank + 1])
    if e > s:
        indices[s:e].copy_(torch.from_numpy(np.ascontiguousarray(X.indices[s:e], dtype=np.int32)))
        data[s:e].copy_(torch.from_numpy(np.ascontiguousarray(X.data[s:e], dtype=np.float32)))
    for r in range(world):
        rs, re = int(eb[r]), int(eb[r + 1])
        if re > rs:
            dist.broadcast(indices[rs:re], src=r)
            dist.broadcast(data[rs:re], src=r)
    return indptr, indices[:nnz], data[:nnz]


def prepare_sharded(adata, lr_df, number_nearest_neighbors: int = 10, use_rep_spatial: str = "X_spatial", **kw):
    """``prepareLRInteraction`` over this rank's block of LR pairs.

    Every rank builds the (replicated) graph and scores only its own pairs; the
    returned AnnData holds the n x P_local block plus the shard record.  The expression
    matrix reaches the GPUs through ``upload_csr_replicated`` (1/world of it per PCIe link)."""
    from .tools import prepareLRInteraction
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    P = len(lr_df)
    names = (lr_df["ligand"].astype(str) + "::" + lr_df["receptor"].astype(str)).to_numpy(dtype=object)
    sh = Shard(rank, world, P, names)
    if "csr_dev" not in kw and world > 1 and dist.get_backend() == "nccl":
        kw["csr_dev"] = upload_csr_replicated(adata.X)
    local = prepareLRInteraction(adata, lr_df.iloc[sh.start:sh.stop].reset_index(drop=True),
                                 number_nearest_neighbors, use_rep_spatial, **kw)
    local.uns[SHARD_KEY] = sh
    return local


def shard_of(lr_adata):
    sh = lr_adata.uns.get(SHARD_KEY) if hasattr(lr_adata, "uns") else None
    return sh if isinstance(sh, Shard) else None
