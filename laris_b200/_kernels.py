"""Torch-tensor carriers around the C ABI.  Tensors only provide device memory
and the stream; every arithmetic step is a call into liblaris_b200.so."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------ stage timer ----
# bench.py brackets the C-ABI calls of a step with CUDA events on the launching stream (``with timing() as t``) to get
# each stage's device time INSIDE the timed region; outside such a block the wrappers add nothing.
class StageTimer:
    def __init__(self):
        self.events = {}
        self.timeline = []          # (stage, start event, end event, host time at enqueue start / end) in call order

    def summary(self):
        """{stage: (calls, total ms)}; call after a synchronisation."""
        return {k: (len(v), float(sum(a.elapsed_time(b) for a, b in v))) for k, v in self.events.items()}


_TIMER = None


class timing:
    def __enter__(self):
        global _TIMER
        self.prev, _TIMER = _TIMER, StageTimer()
        return _TIMER

    def __exit__(self, *exc):
        global _TIMER
        _TIMER = self.prev


def _staged(name):
    def deco(fn):
        def wrapped(*a, **kw):
            t = _TIMER
            if t is None:
                return fn(*a, **kw)
            import time
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0 = time.perf_counter()
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            stage = name(*a, **kw) if callable(name) else name
            t.events.setdefault(stage, []).append((e0, e1))
            t.timeline.append((stage, e0, e1, h0, time.perf_counter()))
            return out
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def _ptr(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("laris_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


PIN_STAGE_MAX_BYTES = 64 << 20
PAGEABLE_PIPELINE = True            # large pageable arrays go up through a ring of page-locked buffers (below)
_RING_SLOTS, _RING_BYTES = 6, 32 << 20
_RING: dict = {}


def _upload_pageable(t, dev):
    """A large PAGEABLE host tensor -> device tensor of the same dtype through a ring of page-locked staging buffers:
    worker threads copy 32 MB pieces into the ring (a plain memcpy, the GIL is released) while the pieces already staged
    go up as asynchronous copies on the current stream.  ``cudaMemcpy`` from pageable memory stages through the driver's
    own bounce buffer on ONE thread (~6-10 GB/s); several staging threads keep the PCIe link (55 GB/s) far busier.
    Returns None when no worker threads are available (the caller then takes the plain path)."""
    from . import _threads
    pool = _threads.pool()
    if pool is None:
        return None
    ring = _RING.get(dev.index)
    if ring is None:
        ring = _RING[dev.index] = [torch.empty(_RING_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(_RING_SLOTS)]
    src = t.reshape(-1).view(torch.uint8)
    out = torch.empty(src.shape, dtype=torch.uint8, device=dev)
    total = src.numel()
    pieces = [(lo, min(total, lo + _RING_BYTES)) for lo in range(0, total, _RING_BYTES)]
    freed = [None] * _RING_SLOTS                              # event after the upload that last used the slot

    def stage(c):
        slot, (lo, hi) = c % _RING_SLOTS, pieces[c]
        if freed[slot] is not None:
            freed[slot].synchronize()                         # the staging buffer is free again
        ring[slot][: hi - lo].copy_(src[lo:hi])
        return slot

    pending = [pool.submit(stage, c) for c in range(min(_RING_SLOTS, len(pieces)))]
    for c, (lo, hi) in enumerate(pieces):
        slot = pending[c].result()
        out[lo:hi].copy_(ring[slot][: hi - lo], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        freed[slot] = ev
        if c + _RING_SLOTS < len(pieces):
            pending.append(pool.submit(stage, c + _RING_SLOTS))
    # the ring is reused by the next call: its last uploads must have left the buffers before anyone refills them
    for ev in freed:
        if ev is not None:
            ev.synchronize()
    return out.view(t.dtype).view(t.shape)


def to_device(a, dtype):
    """numpy -> device tensor of ``dtype``, without blocking the host: a pageable source is first copied into a page-locked
    staging buffer (torch's caching pinned allocator) so that the upload is an asynchronous copy on the current stream -
    a synchronous ``cudaMemcpy`` would make the host wait for every kernel queued before it.  Arrays above 64 MB (the
    expression matrix, which has its own copy stream) go through ``_upload_pageable``'s staging ring."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a) if a.dtype != np.dtype(object) else None
    if t is None:
        raise TypeError("object arrays cannot be uploaded")
    dev = require_cuda()
    if not t.is_pinned() and 0 < t.numel() * t.element_size() <= PIN_STAGE_MAX_BYTES:
        t = t.pin_memory()
    if t.is_pinned():
        out = t.to(dev, non_blocking=True)
        return out if out.dtype == dtype else out.to(dtype)
    if PAGEABLE_PIPELINE and t.numel() > 0:
        out = _upload_pageable(t, dev)
        if out is not None:
            return out if out.dtype == dtype else out.to(dtype)
    return t.to(dev, dtype=dtype)


def to_host(*tensors):
    """Device tensors -> numpy arrays through page-locked staging (one stream sync for all).

    The arrays are views of torch pinned-host tensors: torch's caching host allocator
    recycles the pages once the arrays are garbage collected, so steady-state calls pay
    neither page-locking nor page-fault cost and the copies run at PCIe speed."""
    outs = []
    for t in tensors:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        outs.append(h)
    torch.cuda.current_stream().synchronize()
    return tuple(h.numpy() for h in outs) if len(outs) != 1 else outs[0].numpy()


def to_host_async(*tensors):
    """Queue device -> pinned-host copies of small result tensors NOW and return ``get()``, which waits for exactly these
    copies (an event) and gives the numpy arrays.  Unlike ``t.cpu()`` - a copy queued at the time of the call, behind
    whatever was launched in between - the data is on the host as soon as the kernels before this point are done."""
    outs = []
    for t in tensors:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        outs.append(h)
    done = torch.cuda.Event()
    done.record()

    def get():
        done.synchronize()
        return tuple(h.numpy() for h in outs) if len(outs) != 1 else outs[0].numpy()
    return get


# ----------------------------------------------------------------- K1 ----
@_staged("K1 knn_graph")
def knn_graph(xy, k, include_self, want_order=True):
    n, dim = xy.shape
    dev = xy.device
    idx = torch.empty((n, k), dtype=torch.int32, device=dev)
    dist = torch.empty((n, k), dtype=torch.float64, device=dev)
    order = torch.empty(n, dtype=torch.int32, device=dev) if want_order else None
    check(lib.laris_knn_graph(_ptr(xy, torch.float64), n, dim, k, int(bool(include_self)), _ptr(idx), _ptr(dist),
                              _ptr(order), _stream()), "laris_knn_graph")
    return idx, dist, order


@_staged("K1 decay_weights")
def decay_weights(dist, sigma=None, want64=False):
    m = dist.numel()
    w32 = torch.empty(dist.shape, dtype=torch.float32, device=dist.device)
    w64 = torch.empty(dist.shape, dtype=torch.float64, device=dist.device) if want64 else None
    scale = torch.empty(1, dtype=torch.float64, device=dist.device)
    check(lib.laris_decay_weights(_ptr(dist, torch.float64), m, float(sigma) if sigma else 0.0, _ptr(w64), _ptr(w32),
                                  _ptr(scale), _stream()), "laris_decay_weights")
    return w32, w64, scale


@_staged("K1r radius_graph")
def radius_graph(xy, radius, include_self, want_order=True):
    """K1r: (indptr int64[n+1], indices int32[nnz] column-sorted, dist fp64[nnz], order int32[n])."""
    n, dim = xy.shape
    dev = xy.device
    counts = torch.empty(n, dtype=torch.int64, device=dev)
    order = torch.empty(n, dtype=torch.int32, device=dev) if want_order else None
    check(lib.laris_radius_graph_count(_ptr(xy, torch.float64), n, dim, float(radius), int(bool(include_self)), _ptr(counts),
                                       _ptr(order), _stream()), "laris_radius_graph_count")
    indptr = exclusive_scan(counts)
    nnz = int(indptr[-1].item())
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    dist = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)
    check(lib.laris_radius_graph_fill(_ptr(xy, torch.float64), n, dim, float(radius), int(bool(include_self)), _ptr(indptr),
                                      _ptr(indices), _ptr(dist), _stream()), "laris_radius_graph_fill")
    return indptr, indices[:nnz], dist[:nnz], order


def csr_pad_rows(indptr, indices, w, kmax):
    """Fixed-degree rows (idx int32 [n,kmax], w fp32 [n,kmax]); padding = (own row, weight 0)."""
    n = indptr.numel() - 1
    idx = torch.empty((n, kmax), dtype=torch.int32, device=indptr.device)
    wout = torch.empty((n, kmax), dtype=torch.float32, device=indptr.device)
    check(lib.laris_csr_pad_rows(_ptr(indptr, torch.int64), indices.data_ptr(), w.data_ptr(), n, int(kmax), _ptr(idx),
                                 _ptr(wout), _stream()), "laris_csr_pad_rows")
    return idx, wout


# ------------------------------------------------------ gene selection ----
def exclusive_scan(counts):
    out = torch.empty(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    check(lib.laris_exclusive_scan_i64(_ptr(counts, torch.int64), counts.numel(), _ptr(out), _stream()),
          "laris_exclusive_scan_i64")
    return out


@_staged("select genes")
def csr_select(indptr, indices, data, gene_map):
    """(indptr_u int64[n+1], packed int64[nnz_u]) restricted to gene_map >= 0."""
    n = indptr.numel() - 1
    counts = torch.empty(n, dtype=torch.int64, device=indptr.device)
    check(lib.laris_csr_select_count(_ptr(indptr, torch.int64), _ptr(indices, torch.int32), _ptr(data, torch.float32), n,
                                     _ptr(gene_map, torch.int32), _ptr(counts), _stream()), "laris_csr_select_count")
    out_indptr = exclusive_scan(counts)
    nnz = int(out_indptr[-1].item())
    packed = torch.empty(max(nnz, 1), dtype=torch.int64, device=indptr.device)
    check(lib.laris_csr_select_fill(_ptr(indptr), _ptr(indices), _ptr(data), n, _ptr(gene_map), _ptr(out_indptr),
                                    _ptr(packed), _stream()), "laris_csr_select_fill")
    return out_indptr, packed[:nnz] if nnz else packed[:0]


@_staged("select genes")
def csr_select_rows(indptr, indices, data, gene_map):
    """Single-pass selection: (packed int64 [nnz(X)], row_end int64 [n]); row i's kept entries are
    packed[indptr[i] : row_end[i]].  No count pass, no scan, no host sync; costs nnz(X) slots."""
    n = indptr.numel() - 1
    packed = torch.empty(max(int(indices.numel()), 1), dtype=torch.int64, device=indptr.device)
    row_end = torch.empty(max(n, 1), dtype=torch.int64, device=indptr.device)
    check(lib.laris_csr_select_rows(_ptr(indptr, torch.int64), _ptr(indices, torch.int32), _ptr(data, torch.float32), n,
                                    _ptr(gene_map, torch.int32), _ptr(packed), _ptr(row_end), _stream()),
          "laris_csr_select_rows")
    return packed, row_end


# -------------------------------------------------------------- K2+K3 ----
def lr_column_layout(lig_u, rec_u, n_genes):
    """Host call: (column_of [n_genes] int32, n_columns)."""
    import ctypes
    cols = np.empty(n_genes, dtype=np.int32)
    ncol = ctypes.c_int32(0)
    check(lib.laris_lr_column_layout(lig_u.ctypes.data, rec_u.ctypes.data, len(lig_u), n_genes, cols.ctypes.data,
                                     ctypes.addressof(ncol)), "laris_lr_column_layout")
    return cols, int(ncol.value)


def padded_ld(P):
    """Row stride of the dense score matrix: a multiple of 32 floats, so every 32-column block of a row is one aligned
    128-byte line (the column-blocked null pass keeps such blocks L2-resident; a row piece that straddles two lines
    would cost two tags)."""
    return (P + 31) // 32 * 32


def _cx_args(cx):
    """(ptr, cols, vcol, n_complex, rule) C arguments of a complex table (device int32 tensors) or five nulls."""
    if cx is None:
        return None, None, None, 0, 0
    ptr, cols, vcol, rule = cx
    return _ptr(ptr, torch.int32), _ptr(cols, torch.int32), _ptr(vcol, torch.int32), int(vcol.numel()), int(rule)


@_staged("K2+K3 diffuse_score (sparse)")
def diffuse_lr_score(xu_indptr, xu_packed, G_u, idx, w, order, lig, rec, want_row_nnz=True, cx=None, row_end=None):
    """``row_end=None``: compact rows (row i = packed[indptr[i]:indptr[i+1]]); else row i = packed[indptr[i]:row_end[i]]."""
    n, k = idx.shape
    P = lig.numel()
    ld = padded_ld(P)
    S = torch.empty((n, ld), dtype=torch.float32, device=idx.device)
    row_nnz = torch.empty(n, dtype=torch.int64, device=idx.device) if want_row_nnz else None
    packed_ptr = xu_packed.data_ptr() if xu_packed.numel() else _ptr(torch.zeros(1, dtype=torch.int64, device=idx.device))
    start = _ptr(xu_indptr, torch.int64)
    end = start + 8 if row_end is None else _ptr(row_end, torch.int64)
    check(lib.laris_diffuse_lr_score_rows(start, end, packed_ptr, n, G_u, _ptr(idx, torch.int32),
                                          _ptr(w, torch.float32), k, _ptr(order, torch.int32) if order is not None else None,
                                          _ptr(lig, torch.int32), _ptr(rec, torch.int32), P, _ptr(S), ld, _ptr(row_nnz),
                                          *_cx_args(cx), _stream()), "laris_diffuse_lr_score")
    return S, row_nnz


@_staged("select genes")
def csr_select_dense(indptr, indices, data, gene_map, n_columns):
    """(Xd fp32 [n, ldX] dense over the packed columns, neg_flag int32[1]); ldX = 128-multiple > n_columns."""
    n = indptr.numel() - 1
    ldX = (int(n_columns) + 1 + 127) // 128 * 128
    Xd = torch.empty((n, ldX), dtype=torch.float32, device=indptr.device)
    neg = torch.empty(1, dtype=torch.int32, device=indptr.device)
    check(lib.laris_csr_select_dense(_ptr(indptr, torch.int64), _ptr(indices, torch.int32), _ptr(data, torch.float32), n,
                                     _ptr(gene_map, torch.int32), _ptr(Xd), ldX, _ptr(neg), _stream()),
          "laris_csr_select_dense")
    return Xd, neg


@_staged("K2+K3 diffuse_score (dense)")
def diffuse_lr_score_dense(Xd, neg_flag, G_u, idx, w, order, lig, rec, want_row_nnz=True, cx=None):
    n, k = idx.shape
    P = lig.numel()
    ld = padded_ld(P)
    S = torch.empty((n, ld), dtype=torch.float32, device=idx.device)
    row_nnz = torch.empty(n, dtype=torch.int64, device=idx.device) if want_row_nnz else None
    check(lib.laris_diffuse_lr_score_dense_complex(
        _ptr(Xd, torch.float32), Xd.shape[1], _ptr(neg_flag, torch.int32) if neg_flag is not None else None,
        n, G_u, _ptr(idx, torch.int32), _ptr(w, torch.float32), k, _ptr(order, torch.int32) if order is not None else None,
        _ptr(lig, torch.int32), _ptr(rec, torch.int32), P, _ptr(S), ld, _ptr(row_nnz), *_cx_args(cx), _stream()),
        "laris_diffuse_lr_score_dense")
    return S, row_nnz


@_staged("csr_compact")
def csr_compact(S, P, row_nnz):
    n, ld = S.shape
    indptr = exclusive_scan(row_nnz)
    nnz = int(indptr[-1].item())
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=S.device)
    data = torch.empty(max(nnz, 1), dtype=torch.float32, device=S.device)
    check(lib.laris_csr_compact(_ptr(S, torch.float32), ld, n, P, _ptr(indptr), _ptr(indices), _ptr(data), _stream()),
          "laris_csr_compact")
    return indptr, indices[:nnz], data[:nnz]


def csr_to_dense(indptr, indices, data, n, P):
    ld = padded_ld(P)
    S = torch.empty((n, ld), dtype=torch.float32, device=indptr.device)
    check(lib.laris_csr_to_dense(_ptr(indptr, torch.int64), _ptr(indices, torch.int32) if indices.numel() else None,
                                 _ptr(data, torch.float32) if data.numel() else None, n, P, _ptr(S), ld, _stream()),
          "laris_csr_to_dense")
    return S


# ----------------------------------------------------------------- K4 ----
@_staged("K4 spatial_stats (legacy)")
def spatial_stats(S, P, idx, w, l1_normalise, order=None):
    n, ld = S.shape
    k = idx.shape[1]
    out = torch.empty((5, P), dtype=torch.float64, device=S.device)
    check(lib.laris_spatial_stats(_ptr(S, torch.float32), ld, n, P, _ptr(idx, torch.int32), _ptr(w, torch.float32), k,
                                  int(bool(l1_normalise)), _ptr(order, torch.int32) if order is not None else None,
                                  _ptr(out), _stream()), "laris_spatial_stats")
    return out


@_staged(lambda S, P, idx, w, order, null_idx=None, null_w=None: "K4a observed_stats" if null_idx is None else ("K4b null_stats" if idx is None else "K4a+K4b"))
def spatial_stats_batched(S, P, idx, w, order, null_idx=None, null_w=None):
    """[(1 if idx is given) + R, 5, P] fp64: observed graph (row 0) and the R random graphs ``null_idx`` / ``null_w``
    [R, n, k]; either part may be omitted."""
    n, ld = S.shape
    k = idx.shape[1] if idx is not None else null_idx.shape[2]
    R = 0 if null_idx is None else int(null_idx.shape[0])
    out = torch.empty(((1 if idx is not None else 0) + R, 5, P), dtype=torch.float64, device=S.device)
    check(lib.laris_spatial_stats_batched(_ptr(S, torch.float32), ld, n, P,
                                          _ptr(idx, torch.int32) if idx is not None else None,
                                          _ptr(w, torch.float32) if idx is not None else None,
                                          k, _ptr(order, torch.int32) if order is not None else None,
                                          _ptr(null_idx, torch.int32) if R else None,
                                          _ptr(null_w, torch.float32) if R else None, R, _ptr(out), _stream()),
          "laris_spatial_stats_batched")
    return out


@_staged("K4 pack scores")
def scores_pack(S, P, sp_ptr, nnz):
    """Packed rows of the score matrix: int64 entries {column, fp32 bits}, row i = entries[sp_ptr[i] : sp_ptr[i+1]] - its
    non-zeros, padded with zero-valued entries to a multiple of 32 (``sp_ptr`` = scan of the padded row counts,
    ``nnz`` = sp_ptr[n], or an estimate of it: rows that do not fit are left out)."""
    n, ld = S.shape
    entries = torch.empty(max(int(nnz), 1), dtype=torch.int64, device=S.device)
    check(lib.laris_scores_pack(_ptr(S, torch.float32), ld, n, P, _ptr(sp_ptr, torch.int64), _ptr(entries), entries.numel(), _stream()),
          "laris_scores_pack")
    return entries


def spatial_stats_sparse_max_ld():
    return int(lib.laris_spatial_stats_sparse_max_ld())


@_staged("K4 sparse_stats")
def spatial_stats_sparse(S, P, sp_ptr, entries, idx, w, order, null_idx=None, null_w=None):
    """``spatial_stats_batched`` from the packed rows (``scores_pack``) - for mostly-zero score matrices."""
    n, ld = S.shape
    k = idx.shape[1] if idx is not None else null_idx.shape[2]
    R = 0 if null_idx is None else int(null_idx.shape[0])
    out = torch.empty(((1 if idx is not None else 0) + R, 5, P), dtype=torch.float64, device=S.device)
    check(lib.laris_spatial_stats_sparse(_ptr(S, torch.float32), ld, n, P, _ptr(sp_ptr, torch.int64), _ptr(entries, torch.int64), entries.numel(),
                                         _ptr(idx, torch.int32) if idx is not None else None,
                                         _ptr(w, torch.float32) if idx is not None else None,
                                         k, _ptr(order, torch.int32) if order is not None else None,
                                         _ptr(null_idx, torch.int32) if R else None,
                                         _ptr(null_w, torch.float32) if R else None, R, _ptr(out), _stream()),
          "laris_spatial_stats_sparse")
    return out


@_staged("K4b null_graph")
def null_graph_philox(n, k, w, repeat, seed, out=None):
    cols, wout = out if out is not None else (torch.empty((n, k), dtype=torch.int32, device=w.device),
                                              torch.empty((n, k), dtype=torch.float32, device=w.device))
    check(lib.laris_null_graph_philox(n, k, _ptr(w, torch.float32), int(repeat), int(seed) & (2 ** 64 - 1), _ptr(cols),
                                      _ptr(wout), _stream()), "laris_null_graph_philox")
    return cols, wout


def gather_f32(w, perm):
    out = torch.empty(perm.shape, dtype=torch.float32, device=w.device)
    check(lib.laris_gather_f32(_ptr(w, torch.float32), _ptr(perm, torch.int64), perm.numel(), _ptr(out), _stream()),
          "laris_gather_f32")
    return out


# ----------------------------------------------------------------- K5 ----
@_staged("K5 composition")
def neighborhood_composition(labels, idx, w, C):
    n, k = idx.shape
    N = torch.empty((n, C), dtype=torch.float32, device=idx.device)
    check(lib.laris_neighborhood_composition(_ptr(labels, torch.int32), _ptr(idx, torch.int32), _ptr(w, torch.float32),
                                             n, k, C, _ptr(N), _stream()), "laris_neighborhood_composition")
    return N


MMA_MIN_C, MMA_MAX_C = 1, 64      # cell-type counts the tcgen05 contraction handles (N row = two values per lane)
SPARSE_MAX_ENTRIES_PER_CELL = 8.0   # above this many (a <= b) keys per cell the dense tensor-core contraction wins


def celltype_pair_keys(N):
    """Queue the key histogram of the key-gather contraction (no host synchronisation): int32 [C*C + 1] device tensor to
    hand to ``celltype_pair_contract(key_count=...)`` later."""
    n, C = N.shape
    key_count = torch.empty(C * C + 1, dtype=torch.int32, device=N.device)
    check(lib.laris_celltype_pair_keys(_ptr(N, torch.float32), n, C, _ptr(key_count), _stream()), "laris_celltype_pair_keys")
    return key_count


@_staged("K5 pair_contract")
def celltype_pair_contract(S, P, N, impl=None, order=None, key_count=None):
    """Numerators num[P, C*C] and profile norms qn2[C*C] of the cell-type-pair cosines.
    impl: 0 tcgen05 tensor cores (dense B operand), 1 fp32 CUDA cores, 2 key-gather ("sparse B", fp64 accumulation);
    None = 2 when the neighbourhoods hold few cell types (<= SPARSE_MAX_ENTRIES_PER_CELL keys per cell on average - the
    regime of every kNN graph with k ~ 10), else 0 whenever the number of cell types allows it, else 1."""
    n, ld = S.shape
    C = N.shape[1]
    num = torch.empty((P, C * C), dtype=torch.float64, device=S.device)
    qn2 = torch.empty(C * C, dtype=torch.float64, device=S.device)
    if impl in (None, 2) and C <= MMA_MAX_C:
        if key_count is None:
            key_count = celltype_pair_keys(N)
        total = int(key_count[: C * C].sum().item())             # sizes the entry list and picks the implementation
        if impl == 2 or total <= SPARSE_MAX_ENTRIES_PER_CELL * n:
            check(lib.laris_pair_profile_norm(_ptr(N, torch.float32), n, C, _ptr(qn2), _stream()), "laris_pair_profile_norm")
            check(lib.laris_celltype_pair_contract_sparse(_ptr(S, torch.float32), ld, n, P, _ptr(N, torch.float32), C,
                                                          _ptr(key_count), total, _ptr(num), _stream()),
                  "laris_celltype_pair_contract_sparse")
            return num, qn2
    if impl in (None, 2):
        impl = 0 if MMA_MIN_C <= C <= MMA_MAX_C else 1
    check(lib.laris_celltype_pair_contract(_ptr(S, torch.float32), ld, n, P, _ptr(N, torch.float32), C,
                                           _ptr(order, torch.int32) if order is not None else None, _ptr(num),
                                           _ptr(qn2), int(impl), _stream()), "laris_celltype_pair_contract")
    return num, qn2


@_staged("K5b nonzero_count")
def celltype_nonzero_count(S, P, labels, C):
    n, ld = S.shape
    cnt = torch.empty((C, P), dtype=torch.int64, device=S.device)
    check(lib.laris_celltype_nonzero_count(_ptr(S, torch.float32), ld, n, P, _ptr(labels, torch.int32), C, _ptr(cnt),
                                           _stream()), "laris_celltype_nonzero_count")
    return cnt


@_staged("K5b nonzero_count (packed)")
def celltype_nonzero_count_packed(sp_ptr, entries, P, labels, C):
    """K5b from the packed rows of S (``scores_pack``): [C, P] int64 counts."""
    n = sp_ptr.numel() - 1
    cnt = torch.empty((C, P), dtype=torch.int64, device=sp_ptr.device)
    check(lib.laris_celltype_nonzero_count_packed(_ptr(sp_ptr, torch.int64), _ptr(entries, torch.int64), entries.numel(), n, P,
                                                  _ptr(labels, torch.int32), C, _ptr(cnt), _stream()),
          "laris_celltype_nonzero_count_packed")
    return cnt


@_staged("COSG onehot_contract")
def onehot_contract_csr(xu_indptr, xu_packed, G_u, labels, C, row_end=None):
    n = xu_indptr.numel() - 1
    dev = xu_indptr.device
    s = torch.empty((C, G_u), dtype=torch.float64, device=dev)
    c = torch.empty((C, G_u), dtype=torch.float64, device=dev)
    q = torch.empty(G_u, dtype=torch.float64, device=dev)
    packed_ptr = xu_packed.data_ptr() if xu_packed.numel() else None
    start = _ptr(xu_indptr, torch.int64)
    end = start + 8 if row_end is None else _ptr(row_end, torch.int64)
    check(lib.laris_onehot_contract_rows(start, end, packed_ptr, n, G_u, _ptr(labels, torch.int32), C,
                                         _ptr(s), _ptr(c), _ptr(q), _stream()), "laris_onehot_contract_csr")
    return s, c, q


@_staged("row_qc")
def row_qc(S, P, top_n=()):
    """Per-cell (total fp64 [n], nnz int64 [n], top_sum fp64 [n, len(top_n)]) of the score matrix."""
    n, ld = S.shape
    tops = np.ascontiguousarray(top_n, dtype=np.int32)
    total = torch.empty(n, dtype=torch.float64, device=S.device)
    nnz = torch.empty(n, dtype=torch.int64, device=S.device)
    top = torch.empty((n, max(len(tops), 1)), dtype=torch.float64, device=S.device)
    check(lib.laris_row_qc(_ptr(S, torch.float32), ld, n, P, tops.ctypes.data if len(tops) else None, len(tops),
                           _ptr(total), _ptr(nnz), _ptr(top), _stream()), "laris_row_qc")
    return total, nnz, top[:, :len(tops)]


# -------------------------------------------------------------- K6/K7 ----
@_staged("K6 pvalue_counts")
def pvalue_counts(table, bg, n_perm, *, active=None, seed=0, draws=None, slot_gid=None, P_global=None, g0=0, t0=0,
                  out=None, only_ge=False):
    """Exceedance counts (ge, nz, ge_nz) - or (ge, None, None) with ``only_ge`` (the standard p-value needs no more)."""
    Gn, ldt = table.shape
    P, nb = bg.shape
    dev = table.device
    if out is None:
        out = tuple(torch.zeros((Gn, P), dtype=torch.int32, device=dev) if (i == 0 or not only_ge) else None
                    for i in range(3))
    ge, nz, both = out
    check(lib.laris_pvalue_counts(_ptr(table, torch.float64), ldt, Gn, P, _ptr(bg, torch.int32), nb,
                                  _ptr(active, torch.uint8) if active is not None else None,
                                  _ptr(slot_gid, torch.int64) if slot_gid is not None else None,
                                  int(P if P_global is None else P_global), int(g0), int(t0), int(n_perm),
                                  0 if draws is None else 1, int(seed) & (2 ** 64 - 1),
                                  _ptr(draws, torch.uint8) if draws is not None else None,
                                  _ptr(ge), _ptr(nz) if nz is not None else None,
                                  _ptr(both) if both is not None else None, _stream()), "laris_pvalue_counts")
    return ge, nz, both


@_staged("K7 bh_fdr")
def bh_fdr(p, sel):
    Gn, L = p.shape
    out = torch.empty((Gn, L), dtype=torch.float64, device=p.device)
    check(lib.laris_bh_fdr(_ptr(p, torch.float64), _ptr(sel, torch.uint8), Gn, L, _ptr(out), _stream()), "laris_bh_fdr")
    return out


# --------------------------------------------------- step-2 tables (T1..T5) ----
@_staged("tables")
def pair_cosine_regularise(num, ss, qn2, mu, want_cos=False):
    """(cos or None, reg) [P, L] fp64 from the K5 numerators, the column sums of squares and the profile norms."""
    P, L = num.shape
    cos = torch.empty((P, L), dtype=torch.float64, device=num.device) if want_cos else None
    reg = torch.empty((P, L), dtype=torch.float64, device=num.device)
    check(lib.laris_pair_cosine_regularise(_ptr(num, torch.float64), _ptr(ss, torch.float64), _ptr(qn2, torch.float64),
                                           P, L, float(mu), _ptr(cos), _ptr(reg), _stream()),
          "laris_pair_cosine_regularise")
    return cos, reg


@_staged("tables")
def column_iqr_log(table, q_upper=0.95, q_lower=0.75):
    rows, cols = table.shape
    out = torch.empty_like(table)
    check(lib.laris_column_iqr_log(_ptr(table, torch.float64), rows, cols, float(q_upper), float(q_lower), _ptr(out),
                                   _stream()), "laris_column_iqr_log")
    return out


@_staged("tables")
def interaction_scores(spec_l, spec_r, factor, cnt, n_c, pair_id, P):
    Pt, C = spec_l.shape
    inter = torch.empty((Pt, C * C), dtype=torch.float64, device=spec_l.device)
    check(lib.laris_interaction_scores(_ptr(spec_l, torch.float64), _ptr(spec_r, torch.float64),
                                       _ptr(factor, torch.float64), _ptr(cnt, torch.int64), _ptr(n_c, torch.float64),
                                       _ptr(pair_id, torch.int32), Pt, int(P), C, _ptr(inter), _stream()),
          "laris_interaction_scores")
    return inter


@_staged("tables")
def topk_mean(values, k=100):
    """Device pair (mean of the k largest, 0.1 / mean or 1)."""
    out = torch.empty(2, dtype=torch.float64, device=values.device)
    check(lib.laris_topk_mean(_ptr(values, torch.float64), values.numel(), int(k), _ptr(out), _stream()),
          "laris_topk_mean")
    return out


@_staged("tables")
def finalize_scores(inter, rescale2, ctc, pair_id, P, score_threshold):
    Pt, L = inter.shape
    dev = inter.device
    flat = torch.empty((Pt, L), dtype=torch.float64, device=dev)
    table = torch.empty((L, P), dtype=torch.float64, device=dev)
    active = torch.empty(P, dtype=torch.uint8, device=dev)
    check(lib.laris_finalize_scores(_ptr(inter, torch.float64), _ptr(rescale2, torch.float64), _ptr(ctc, torch.float64),
                                    _ptr(pair_id, torch.int32), Pt, int(P), L, float(score_threshold), _ptr(flat),
                                    _ptr(table), _ptr(active), _stream()), "laris_finalize_scores")
    return flat, table, active


@_staged("tables")
def pvalues_from_counts(table, active, ge, nz, both, n_perm, conditional, prefilter, prefilter_threshold):
    L, P = table.shape
    dev = table.device
    p = torch.empty((L, P), dtype=torch.float64, device=dev)
    pr = torch.empty((L, P), dtype=torch.float64, device=dev)
    sel = torch.empty((L, P), dtype=torch.uint8, device=dev)
    check(lib.laris_pvalues_from_counts(_ptr(table, torch.float64), _ptr(active, torch.uint8), _ptr(ge, torch.int32),
                                        _ptr(nz, torch.int32) if nz is not None else None,
                                        _ptr(both, torch.int32) if both is not None else None, L, P, int(n_perm),
                                        int(bool(conditional)), int(bool(prefilter)), float(prefilter_threshold),
                                        _ptr(p), _ptr(pr), _ptr(sel), _stream()), "laris_pvalues_from_counts")
    return p, pr, sel


@_staged("tables")
def gather_result_rows(p_tab, fdr_tab, pair_id):
    L, P = p_tab.shape
    Pt = pair_id.numel()
    dev = p_tab.device
    outs = [torch.empty((Pt, L), dtype=torch.float64, device=dev) for _ in range(3)]
    check(lib.laris_gather_result_rows(_ptr(p_tab, torch.float64), _ptr(fdr_tab, torch.float64), _ptr(pair_id, torch.int32),
                                       Pt, P, L, *[_ptr(o) for o in outs], _stream()), "laris_gather_result_rows")
    return outs


@_staged("tables")
def result_row_order(inter, flat, C):
    """T6a: (order int32 [Pt*C*C], exotic int32 [1]) - the reference's table row order, sorted on the device."""
    total = flat.numel()
    Pt = total // (C * C)
    order = torch.empty(total, dtype=torch.int32, device=flat.device)
    exotic = torch.empty(1, dtype=torch.int32, device=flat.device)
    check(lib.laris_result_row_order(_ptr(inter, torch.float64), _ptr(flat, torch.float64), Pt, C, _ptr(order),
                                     _ptr(exotic), _stream()), "laris_result_row_order")
    return order, exotic


@_staged("tables")
def result_rows_gather(order, C, flat, p_rows=None, fdr_rows=None, nlog_rows=None, extra=None):
    """T6b: the numeric columns of the result table and the per-row (pair, sender, receiver) codes in table order.
    Returns (score, p, fdr, nlog, extra, pair, sender, receiver); columns without a source are None."""
    total = order.numel()
    dev = order.device
    f64 = lambda src: None if src is None else torch.empty(total, dtype=torch.float64, device=dev)   # noqa: E731
    score = torch.empty(total, dtype=torch.float64, device=dev)
    outs = [f64(p_rows), f64(fdr_rows), f64(nlog_rows), f64(extra)]
    codes = [torch.empty(total, dtype=torch.int32, device=dev) for _ in range(3)]
    check(lib.laris_result_rows_gather(_ptr(order, torch.int32), total, C, _ptr(flat, torch.float64),
                                       _ptr(p_rows, torch.float64), _ptr(fdr_rows, torch.float64),
                                       _ptr(nlog_rows, torch.float64), _ptr(extra, torch.float64), _ptr(score),
                                       *[_ptr(o) for o in outs], *[_ptr(c) for c in codes], _stream()),
          "laris_result_rows_gather")
    return (score, *outs, *codes)


# ------------------------------------------- label-permutation null (extension) ----
@_staged("label null: permute")
def permute_labels(labels, permutation, seed):
    out = torch.empty_like(labels)
    check(lib.laris_permute_labels(_ptr(labels, torch.int32), labels.numel(), int(permutation), int(seed) & (2 ** 64 - 1),
                                   _ptr(out), _stream()), "laris_permute_labels")
    return out


def count_exceed(obs, val, ge):
    check(lib.laris_count_exceed(_ptr(obs, torch.float64), _ptr(val, torch.float64), obs.numel(), _ptr(ge, torch.int32),
                                 _stream()), "laris_count_exceed")
    return ge
