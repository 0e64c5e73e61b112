"""Device pipeline of the LR scoring path.

Every function here is host orchestration only: it moves AnnData arrays to
the GPU (torch tensors as the carrier), calls the C-ABI kernels in
``liblaris_b200.so`` and brings small per-pair / per-group tables back.  The
reference functions each stage replaces are cited on the kernels
(include/laris_b200.h) and on the public wrappers in ``laris_b200.tools``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import os
import threading

import numpy as np
import scipy.sparse as sp
import torch

from . import _kernels as K
from . import _rng_compat


# ---------------------------------------------------------------- graph ----
@dataclass
class SpatialGraph:
    idx: torch.Tensor      # [n,k] int32 neighbour ids (CSR indices, indptr = i*k)
    dist: torch.Tensor     # [n,k] fp64 distances
    w: torch.Tensor        # [n,k] fp32 decay weights
    w64: torch.Tensor      # [n,k] fp64 decay weights (the reference's CSR data)
    order: torch.Tensor    # [n] int32 spatial processing order
    scale: torch.Tensor    # [1] fp64 decay scale s
    include_self: bool

    @property
    def n(self):
        return self.idx.shape[0]

    @property
    def k(self):
        return self.idx.shape[1]

    csr: tuple | None = None   # radius graphs: (indptr int64, indices int32, dist fp64, w64 fp64) on the device
    xy: torch.Tensor | None = None   # the coordinates the search ran on (device), for re-use by a later graph
    xy_host: np.ndarray | None = None   # ... and the host array they were uploaded from, if any

    def to_scipy(self):
        """The reference's ``cellxcell`` (n x n CSR, fp64 weights, int64 indices)."""
        n, k = self.idx.shape
        if self.csr is not None:
            indptr, indices, _, w64 = self.csr
            return sp.csr_matrix((w64.cpu().numpy(), indices.cpu().numpy().astype(np.int64), indptr.cpu().numpy()),
                                 shape=(n, n))
        return sp.csr_matrix((self.w64.cpu().numpy().ravel(), self.idx.cpu().numpy().astype(np.int64).ravel(),
                              np.arange(0, n * k + 1, k, dtype=np.int64)), shape=(n, n))


def _coords_to_device(xy):
    """Coordinates as an fp64 device tensor.  Host arrays are checked for NaN / inf here: the device graph build is
    asynchronous and can only flag non-finite coordinates in its outputs (index -1)."""
    if torch.is_tensor(xy):
        xy_d = xy
    else:
        xy = np.asarray(xy, dtype=np.float64)
        if xy.ndim == 2 and not np.isfinite(xy).all():
            raise ValueError("spatial coordinates contain NaN or infinite values")
        xy_d = K.to_device(xy, torch.float64)
    if xy_d.dim() != 2 or xy_d.shape[1] not in (2, 3):
        raise ValueError(f"spatial coordinates must be n x 2 or n x 3, got {tuple(xy_d.shape)}")
    return xy_d


def _check_sigma(sigma):
    """The reference divides by sigma (laris/tools/_utils.py:392): zero / negative / non-finite scales are rejected up
    front (the C ABI uses sigma <= 0 as its own marker for the mean(d)/2 rule)."""
    if sigma is not None and not (np.isfinite(sigma) and sigma > 0):
        raise ValueError(f"sigma must be a positive finite number (got {sigma!r}); pass None for the mean(d)/2 rule")


def build_graph(xy, k, include_self, sigma=None, like: SpatialGraph | None = None) -> SpatialGraph:
    """kNN graph + decay weights (K1).  ``sigma=None`` -> s = mean(d)/2.  ``like``: a graph built earlier in the same
    run; when it was searched with the same k / self rule over the SAME coordinates (compared on the device), its
    neighbour lists are re-used and only the weights are recomputed (runLARIS builds two graphs that differ only in
    the decay rule, reference laris/tools/_utils.py:383-394 and :846-853)."""
    _check_sigma(sigma)
    xy_host = None if torch.is_tensor(xy) else np.asarray(xy)
    same = (like is not None and like.xy is not None and like.csr is None and like.k == int(k)
            and like.include_self == bool(include_self))
    if same:
        # the coordinates are compared WITHOUT a device round trip (a torch.equal would make the host wait for everything
        # queued so far): the same device tensor, or host arrays with equal contents
        if torch.is_tensor(xy):
            same = like.xy.data_ptr() == xy.data_ptr() and tuple(like.xy.shape) == tuple(xy.shape)
        else:
            same = like.xy_host is not None and (like.xy_host is xy_host or (like.xy_host.shape == xy_host.shape
                                                                          and np.array_equal(like.xy_host, xy_host)))
    if same:
        xy_d, idx, dist, order = like.xy, like.idx, like.dist, like.order
    else:
        xy_d = _coords_to_device(xy).contiguous()
        idx, dist, order = K.knn_graph(xy_d, int(k), include_self)
    w32, w64, scale = K.decay_weights(dist, sigma, want64=True)
    return SpatialGraph(idx, dist, w32, w64, order, scale, bool(include_self), xy=xy_d, xy_host=xy_host)


def build_radius_graph(xy, radius, include_self, sigma=None) -> SpatialGraph:
    """Radius graph + decay weights (K1r; extension, oracle sklearn.neighbors.radius_neighbors_graph).

    The variable-degree CSR is kept in ``csr``; ``idx``/``w`` hold the same rows padded to the maximum degree with
    (own row, weight 0) edges, which is the fixed-degree layout every gather kernel takes.  ``sigma=None`` uses
    s = mean(stored distances)/2, the reference's rule for its kNN graphs."""
    _check_sigma(sigma)
    xy_d = _coords_to_device(xy)
    indptr, indices, dist, order = K.radius_graph(xy_d.contiguous(), float(radius), include_self)
    if dist.numel() == 0:
        raise ValueError(f"radius={radius} leaves every cell without a neighbour")
    w32, w64, scale = K.decay_weights(dist, sigma, want64=True)
    kmax = max(int((indptr[1:] - indptr[:-1]).max().item()), 1)
    idx, w = K.csr_pad_rows(indptr, indices, w32, kmax)
    return SpatialGraph(idx, None, w, None, order, scale, bool(include_self), (indptr, indices, dist, w64))


# ----------------------------------------------------------- expression ----
@dataclass
class ExpressionU:
    """Expression restricted to the LR genes, in one of the two device layouts of K2+K3:
    packed sparse rows (``indptr``/``packed``) or dense fp32 rows (``dense``/``neg_flag``)."""
    indptr: torch.Tensor | None   # [n+1] int64
    packed: torch.Tensor | None   # [nnz] int64 {col:int32, val:fp32}
    genes: np.ndarray      # [n_genes] global gene column of each selected gene
    n: int
    columns: np.ndarray | None = None   # [n_genes] packed column of each selected gene (None: 0..n_genes-1)
    n_columns: int | None = None
    dense: torch.Tensor | None = None     # [n, ldX] fp32
    neg_flag: torch.Tensor | None = None  # [1] int32: a negative value was seen
    cx: tuple | None = None               # complex table (ptr, cols, vcol int32 device tensors, rule) or None
    row_end: torch.Tensor | None = None   # [n] int64: rows are packed[indptr[i]:row_end[i]] (single-pass selection);
                                          # None: compact rows packed[indptr[i]:indptr[i+1]]

    @property
    def G_u(self):
        """Column range of the packed matrix (>= number of selected genes when a layout leaves holes)."""
        return len(self.genes) if self.n_columns is None else self.n_columns


def upload_csr(X):
    """scipy sparse -> (indptr int64, indices int32, data fp32) on the device."""
    if not sp.issparse(X):
        raise TypeError("adata.X must be a scipy sparse matrix (as in the reference, laris/tools/__init__.py:100)")
    X = X.tocsr()
    if not X.has_canonical_format:                  # the kernels assume one entry per (row, column)
        X = X.copy()
        X.sum_duplicates()
    return (K.to_device(X.indptr, torch.int64), K.to_device(X.indices, torch.int32), K.to_device(X.data, torch.float32))


_COPY_STREAMS: dict = {}


def upload_csr_async(X):
    """``upload_csr`` on a dedicated copy stream: the host->device copies of the expression matrix (the bulk of the
    input bytes) run while the caller builds the graph on the current stream.  Returns ``(csr_dev, join)``; call
    ``join()`` on the consuming stream before the first kernel that reads the matrix."""
    K.require_cuda()                     # "no CPU fallback": fail loudly before touching a stream
    if not sp.issparse(X):
        raise TypeError("adata.X must be a scipy sparse matrix (as in the reference, laris/tools/__init__.py:100)")
    cur = torch.cuda.current_stream()
    dev = torch.cuda.current_device()
    side = _COPY_STREAMS.get(dev)
    if side is None:
        side = _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        csr_dev = upload_csr(X)
        done = torch.cuda.Event()
        done.record(side)

    def join():
        c = torch.cuda.current_stream()
        c.wait_event(done)
        for t in csr_dev:
            t.record_stream(c)
        return csr_dev
    return csr_dev, join


DENSE_PANEL_MIN_DENSITY = 0.15     # below this the packed sparse rows are the faster layout (measured, profiles/)


def choose_layout(X, layout="auto"):
    """'dense' or 'sparse' device layout of the LR-gene expression for K2+K3.  The fraction of
    non-zeros of the whole matrix stands in for that of the LR genes (no pass over the data)."""
    if layout in ("dense", "sparse"):
        return layout
    if layout != "auto":
        raise ValueError(f"layout must be 'auto', 'dense' or 'sparse', got {layout!r}")
    n, G = X.shape
    return "dense" if X.nnz >= DENSE_PANEL_MIN_DENSITY * float(n) * float(G) else "sparse"


SINGLE_PASS_MAX_BYTES = 24 << 30      # scratch the single-pass selection may take (8 bytes per stored entry of X)


def _select_sparse(indptr, indices, data, gene_map_d, genes, n, cols, n_columns, cx) -> ExpressionU:
    """Packed sparse rows of the LR genes.  Single pass (rows keep their offsets in X, no count / scan / host sync) when
    its nnz(X)-slot scratch is affordable, else the compact two-pass form."""
    if 8 * int(indices.numel()) <= SINGLE_PASS_MAX_BYTES:
        packed, row_end = K.csr_select_rows(indptr, indices, data, gene_map_d)
        return ExpressionU(indptr, packed, genes, n, cols, n_columns, cx=cx, row_end=row_end)
    out_indptr, packed = K.csr_select(indptr, indices, data, gene_map_d)
    return ExpressionU(out_indptr, packed, genes, n, cols, n_columns, cx=cx)


def select_genes(X, genes, csr_dev=None, columns=None, n_columns=None, layout="sparse") -> ExpressionU:
    """Restrict X (n x G CSR) to ``genes`` (unique global columns) on the device.
    ``columns`` optionally gives the packed column of each gene (see ``lr_column_layout``)."""
    n, G = X.shape
    genes = np.asarray(genes, dtype=np.int64)
    gene_map = np.full(G, -1, dtype=np.int32)
    gene_map[genes] = np.arange(len(genes), dtype=np.int32) if columns is None else np.asarray(columns, dtype=np.int32)
    indptr, indices, data = csr_dev if csr_dev is not None else upload_csr(X)
    cols = None if columns is None else np.asarray(columns, dtype=np.int32)
    if layout == "dense":
        ncol = len(genes) if n_columns is None else int(n_columns)
        Xd, neg = K.csr_select_dense(indptr, indices, data, K.to_device(gene_map, torch.int32), ncol)
        return ExpressionU(None, None, genes, n, cols, n_columns, Xd, neg)
    return _select_sparse(indptr, indices, data, K.to_device(gene_map, torch.int32), genes, n, cols, n_columns, None)


def lr_column_layout(lig_u, rec_u, n_genes):
    """Bank-conflict-free numbering of the packed columns for the score kernel
    (host-side graph colouring of the LR table, ``laris_lr_column_layout``)."""
    lig_u = np.ascontiguousarray(lig_u, dtype=np.int32)
    rec_u = np.ascontiguousarray(rec_u, dtype=np.int32)
    return K.lr_column_layout(lig_u, rec_u, int(n_genes))


@dataclass
class LRPlan:
    """Everything about an LR table that does not depend on the cells: the distinct genes, their
    packed (bank-conflict-free) columns and the per-pair column indices, already on the device."""
    genes: np.ndarray
    columns: np.ndarray
    n_columns: int
    lig_c: np.ndarray
    rec_c: np.ndarray
    gene_map_d: torch.Tensor
    lig_d: torch.Tensor
    rec_d: torch.Tensor
    cx: tuple | None = None


_PLAN_CACHE: dict = {}


COMPLEX_RULES = {"min": 0, "gmean": 1}


def plan_lr_table(lig_g, rec_g, G, complexes=None, rule="min") -> LRPlan:
    """Plan of an LR table (cached: the same table is usually scored on many samples).

    ``lig_g`` / ``rec_g`` hold gene columns in [0, G); an entry ``G + c`` names the multi-subunit complex
    ``complexes[c]`` (an array of gene columns), aggregated with ``rule`` ('min' | 'gmean') in the kernel."""
    lig_g = np.ascontiguousarray(lig_g, dtype=np.int64)
    rec_g = np.ascontiguousarray(rec_g, dtype=np.int64)
    if rule not in COMPLEX_RULES:
        raise ValueError(f"complex_rule must be one of {sorted(COMPLEX_RULES)}, got {rule!r}")
    cx_key = None if not complexes else (rule, tuple(tuple(int(g) for g in c) for c in complexes))
    key = (int(G), torch.cuda.current_device(), hash(lig_g.tobytes()), hash(rec_g.tobytes()), len(lig_g), hash(cx_key))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= 8:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        plan = _PLAN_CACHE[key] = _plan_lr_table(lig_g, rec_g, G, complexes or [], rule)
    return plan


def _plan_lr_table(lig_g, rec_g, G, complexes=(), rule="min") -> LRPlan:
    P = len(lig_g)
    sub = np.concatenate([np.asarray(c, dtype=np.int64) for c in complexes]) if len(complexes) else np.empty(0, np.int64)
    if len(complexes) and (min(len(c) for c in complexes) < 1 or sub.min() < 0 or sub.max() >= G):
        raise ValueError("every complex needs at least one subunit gene in [0, G)")
    # units = genes and complexes (pseudo ids >= G); every unit gets a packed column of the D slab
    units, inv = np.unique(np.concatenate([np.asarray(lig_g), np.asarray(rec_g), sub]), return_inverse=True)
    if units[-1] >= G + len(complexes):
        raise ValueError("LR table names a complex that is not in `complexes`")
    lig_u, rec_u = inv[:P].astype(np.int32), inv[P:2 * P].astype(np.int32)
    columns, n_columns = lr_column_layout(lig_u, rec_u, len(units))
    is_gene = units < G
    genes, gene_cols = units[is_gene], columns[is_gene]
    gene_map = np.full(G, -1, dtype=np.int32)
    gene_map[genes] = gene_cols
    lig_c, rec_c = columns[lig_u], columns[rec_u]
    cx = None
    used = units[~is_gene] - G                                    # complexes some pair names
    if len(used):
        ptr = np.concatenate([[0], np.cumsum([len(complexes[c]) for c in used])]).astype(np.int32)
        cols = np.concatenate([gene_map[np.asarray(complexes[c], dtype=np.int64)] for c in used]).astype(np.int32)
        vcol = columns[~is_gene].astype(np.int32)
        cx = (K.to_device(ptr, torch.int32), K.to_device(cols, torch.int32), K.to_device(vcol, torch.int32),
              COMPLEX_RULES[rule])
    return LRPlan(genes, gene_cols, n_columns, lig_c, rec_c, K.to_device(gene_map, torch.int32),
                  K.to_device(lig_c.astype(np.int32), torch.int32), K.to_device(rec_c.astype(np.int32), torch.int32), cx)


def select_planned(plan: LRPlan, csr_dev, n, layout) -> ExpressionU:
    """Device-only gene selection for a planned LR table (no host work, no host sync in the dense layout)."""
    indptr, indices, data = csr_dev
    if layout == "dense":
        Xd, neg = K.csr_select_dense(indptr, indices, data, plan.gene_map_d, plan.n_columns)
        return ExpressionU(None, None, plan.genes, n, plan.columns, plan.n_columns, Xd, neg, plan.cx)
    return _select_sparse(indptr, indices, data, plan.gene_map_d, plan.genes, n, plan.columns, plan.n_columns, plan.cx)


def select_lr_genes(X, lig_g, rec_g, csr_dev=None, layout="auto", complexes=None, rule="min"):
    """Gene selection for an LR table given as global gene columns (entries >= G name ``complexes``).
    Returns (ExpressionU, ligand packed columns [P], receptor packed columns [P]) - the columns as device tensors."""
    plan = plan_lr_table(lig_g, rec_g, X.shape[1], complexes, rule)
    xu = select_planned(plan, csr_dev if csr_dev is not None else upload_csr(X), X.shape[0], choose_layout(X, layout))
    return xu, plan.lig_d, plan.rec_d


# --------------------------------------------------------------- scores ----
@dataclass
class ScoreMatrix:
    S: torch.Tensor                    # [n, ld] fp32 dense, columns >= P are zero
    P: int
    row_nnz: torch.Tensor | None = None
    cache: dict = field(default_factory=dict)

    @property
    def n(self):
        return self.S.shape[0]


def lr_scores(xu: ExpressionU, graph: SpatialGraph, lig_u, rec_u, queue_pointers: bool = True) -> ScoreMatrix:
    """Fused diffusion + LR score (K2+K3).  ``queue_pointers``: also start the row pointers of the packed form and
    their total on the way to the host (not possible while a CUDA graph is being captured)."""
    lig_d = lig_u if torch.is_tensor(lig_u) else K.to_device(np.asarray(lig_u, dtype=np.int32), torch.int32)
    rec_d = rec_u if torch.is_tensor(rec_u) else K.to_device(np.asarray(rec_u, dtype=np.int32), torch.int32)
    if xu.dense is not None:
        S, row_nnz = K.diffuse_lr_score_dense(xu.dense, xu.neg_flag, xu.G_u, graph.idx, graph.w, graph.order, lig_d, rec_d,
                                              cx=xu.cx)
    else:
        S, row_nnz = K.diffuse_lr_score(xu.indptr, xu.packed, xu.G_u, graph.idx, graph.w, graph.order, lig_d, rec_d,
                                        cx=xu.cx, row_end=xu.row_end)
    sm = ScoreMatrix(S, len(lig_u), row_nnz)
    if queue_pointers:
        queue_row_pointers(sm)
    return sm


def queue_row_pointers(sm: "ScoreMatrix"):
    """Row pointers of the packed form of S (exclusive scan of the row counts, rounded up to whole 32-entry chunks) and
    their total, which starts its way to the host now - ``packed_scores`` needs it to size the entry array and by then
    it has arrived."""
    if sm.row_nnz is None:
        sm.row_nnz = K.row_qc(sm.S, sm.P, ())[1].to(torch.int64)
    sm.cache["sp_ptr"] = K.exclusive_scan((sm.row_nnz + 31) & ~31)
    sm.cache["nnz_get"] = K.to_host_async(sm.cache["sp_ptr"][-1:])


def sparse_stats_pays(n, P, k, n_obs, n_null, n_entries):
    """Cost model (ms at 1e6 cells, fitted to B200 measurements, tools/time_k4s.py) choosing between the dense K4 passes
    (tiles for the observed graph, L2-blocked / streaming for the nulls: time ~ P per cell and graph) and the packed-row
    pass (time ~ a latency floor per (cell, graph) + the entries it gathers, + the packing pass).  The packed pass wins for
    wide, mostly-zero tables (config 3: P = 2000 at 10 % non-zeros, 21.7 ms against 37.3 ms) and loses for narrow shards
    or above ~15 % non-zeros."""
    row = n_entries / max(n, 1)
    dense = n * P * (n_null * 4.95e-3 + n_obs * 3.8e-3)
    sparse = n * (n_obs + n_null) * (1.9 + 0.0018 * k * row) + n * P * 1.0e-3
    return k <= 32 and sparse < 0.9 * dense


_ENTRY_HINT: dict = {}            # (n, P) -> packed entries of the last score matrix of that shape


def packed_scores(sm: "ScoreMatrix", k, n_obs, n_null, speculative=False):
    """(sp_ptr, entries) - the packed rows of S for the sparse K4 pass - or None when the dense passes are expected to be
    faster (``sparse_stats_pays``) or S is too wide for the pass.  LARIS_B200_K4_SPARSE=0 / 1 forces the dense / the
    packed pass (A/B runs).

    The entry array is sized by the number of packed entries, a device-side result.  ``speculative``: when a score matrix
    of this shape was packed before, size the array (and take the decision) from THAT count plus 6 % instead of waiting
    for this one - the host then never blocks between the scoring kernels and the statistics pass; the caller must call
    ``packed_ok`` once the statistics are on the host and redo them if it says no (rows beyond the array were skipped)."""
    mode = os.environ.get("LARIS_B200_K4_SPARSE", "auto")
    if mode == "0" or sm.S.shape[1] > K.spatial_stats_sparse_max_ld() or sm.P > 12800 or k > 64:
        return None
    if "packed" in sm.cache:
        return sm.cache["packed"]
    if "nnz_get" not in sm.cache:
        queue_row_pointers(sm)
    hint = _ENTRY_HINT.get((sm.n, sm.P)) if speculative else None
    if hint is None:
        n_entries = capacity = int(np.asarray(sm.cache["nnz_get"]()).ravel()[0])
        _ENTRY_HINT[(sm.n, sm.P)] = n_entries
    else:
        n_entries = hint
        capacity = (int(hint * 1.06) + 1024) // 32 * 32
        sm.cache["packed_check"] = capacity
    if mode != "1" and not sparse_stats_pays(sm.n, sm.P, k, n_obs, n_null, n_entries):
        sm.cache.pop("packed_check", None)
        return None
    packed = (sm.cache["sp_ptr"], K.scores_pack(sm.S, sm.P, sm.cache["sp_ptr"], capacity))
    sm.cache["packed"] = packed
    return packed


def packed_ok(sm: "ScoreMatrix"):
    """False if the speculatively sized entry array of ``packed_scores`` turned out too small for this score matrix
    (its statistics are then invalid and the packed rows are dropped).  Call after a host synchronisation."""
    capacity = sm.cache.pop("packed_check", None)
    if capacity is None:
        return True
    actual = int(np.asarray(sm.cache["nnz_get"]()).ravel()[0])
    _ENTRY_HINT[(sm.n, sm.P)] = actual
    if actual <= capacity:
        return True
    sm.cache.pop("packed", None)
    return False


class CapturedScoring:
    """The device-resident scoring step - graph build, decay weights, gene selection, fused diffusion / LR score -
    captured once as a CUDA graph for a fixed problem shape and replayed with no host launches.

    Every stage of the step is asynchronous (the kNN grid is derived on the device, the single-pass selection needs no
    scan, scratch comes from the stream-ordered allocator), so the ~20 kernels of a step can be recorded into one graph;
    a replay re-runs them on the CURRENT contents of the same device buffers (``xy_d`` and ``csr_dev`` may be
    overwritten in place between replays, e.g. resampled expression values on a fixed sparsity pattern, or the same
    tissue scored against another LR table of the same shape).  Replays are immune to host launch jitter, which
    otherwise paces the many small kernels of the graph build."""

    def __init__(self, xy_d, csr_dev, plan: LRPlan, k: int, layout: str, include_self: bool = True):
        from ._lib import lib
        n = xy_d.shape[0]
        self._run = lambda: self._step(xy_d, csr_dev, plan, n, int(k), layout, include_self)
        self._run()                                    # eager pass: lazy initialisation, function attributes, pools
        torch.cuda.synchronize()
        self.graph_exec = torch.cuda.CUDAGraph()
        before = lib.laris_launch_count()
        with torch.cuda.graph(self.graph_exec):
            self.graph, self.xu, self.scores = self._run()
        self.launches = int(lib.laris_launch_count() - before)       # kernels recorded = kernels run per replay

    @staticmethod
    def _step(xy_d, csr_dev, plan, n, k, layout, include_self):
        graph = build_graph(xy_d, k, include_self, None)
        xu = select_planned(plan, csr_dev, n, layout)
        return graph, xu, lr_scores(xu, graph, plan.lig_d, plan.rec_d, queue_pointers=False)

    def replay(self) -> ScoreMatrix:
        """Run the captured step; the returned ScoreMatrix (and ``self.graph`` / ``self.xu``) live in graph-owned
        buffers that the next replay overwrites."""
        self.graph_exec.replay()
        self.scores.cache.clear()                      # statistics / packed rows of the previous contents
        return self.scores


FP_SAMPLES = 1 << 18          # strided sample of the stored values a content fingerprint covers (every value below that)


def device_fingerprint_async(data_d):
    """Content fingerprint of a device ``.data`` array (see tools/_utils.py): returns a thunk giving (nnz, checksum)
    once the stream work queued so far is complete."""
    nnz = int(data_d.numel())
    if nnz == 0:
        return lambda: (0, 0)
    stride = max(1, nnz // FP_SAMPLES)
    t = data_d[::stride].contiguous().view(torch.int32).sum(dtype=torch.int64).reshape(1)
    h = torch.empty(1, dtype=torch.int64, pin_memory=True)
    h.copy_(t, non_blocking=True)
    done = torch.cuda.Event()
    done.record()

    def result():
        done.synchronize()
        return nnz, int(h.item())
    return result


class PendingCSR(sp.csr_matrix):
    """A scipy CSR matrix whose three arrays may still be on their way from the device.

    ``prepareLRInteraction`` returns the score matrix as soon as the device->host copies are QUEUED (on a copy stream of
    their own): the arrays already exist - views of page-locked buffers - and the first access to ``data`` / ``indices`` /
    ``indptr`` (any scipy operation, ``nnz``, a copy, pickling) waits for the copies' event, once.  A following
    ``runLARIS`` works on the device-resident scores and never touches them, so the 1.6 GB download of a 1e6-cell
    matrix overlaps its kernels instead of preceding them.  Copies, slices and pickles are ordinary matrices."""

    _laris_lock = threading.RLock()          # joins are rare; one lock for all pending matrices keeps instances plain

    def _laris_join(self):
        if "_laris_wait" not in self.__dict__:
            return
        with PendingCSR._laris_lock:           # a second reader waits here until the first one has joined the copies
            wait = self.__dict__.get("_laris_wait")
            if wait is not None:
                wait(self)                     # (touches the arrays through __dict__ / the setters only)
                self.__dict__.pop("_laris_wait", None)

    def _laris_pending(self):
        """True while nobody has looked at the arrays (they then still equal the device copy)."""
        return "_laris_wait" in self.__dict__

    def __reduce__(self):
        return (sp.csr_matrix, ((self.data, self.indices, self.indptr), self.shape))


def _pending_array(name):
    slot = "_laris_" + name

    def get(self):
        self._laris_join()
        return self.__dict__[slot]

    def put(self, value):
        self.__dict__[slot] = value
    return property(get, put)


for _name in ("data", "indices", "indptr"):
    setattr(PendingCSR, _name, _pending_array(_name))

_D2H_STREAMS: dict = {}


def scores_to_csr_async(sm: ScoreMatrix, dtype=np.float32):
    """Start the device CSR compaction and the device->host copies of the result (on a dedicated copy stream, so that
    kernels queued afterwards on the caller's stream overlap them).  Returns ``finish(wait=True)`` giving (scipy matrix,
    content fingerprint): ``wait=False`` returns a ``PendingCSR`` immediately, whose arrays are joined on first access."""
    if sm.row_nnz is None:
        raise ValueError("score matrix has no row counts")
    indptr_d, indices_d, data_d = K.csr_compact(sm.S, sm.P, sm.row_nnz)
    nnz = int(data_d.numel())
    if nnz < 2 ** 31 - 1:                            # scipy's own choice of index dtype for this size, made on the device
        indptr_d = indptr_d.to(torch.int32)
    else:
        indices_d = indices_d.to(torch.int64)
    parts = (indptr_d, indices_d, data_d)
    fp = device_fingerprint_async(data_d)
    cur = torch.cuda.current_stream()
    dev = torch.cuda.current_device()
    side = _D2H_STREAMS.get(dev)
    if side is None:
        side = _D2H_STREAMS[dev] = torch.cuda.Stream(device=dev)
    side.wait_stream(cur)
    host = []
    with torch.cuda.stream(side):
        for t in parts:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            t.record_stream(side)
            host.append(h)
        done = torch.cuda.Event()
        done.record(side)
    np_dtype = np.dtype(dtype)

    def join(X):
        done.synchronize()
        if np_dtype != np.float32:
            X.data = X.__dict__["_laris_data"].astype(np_dtype)

    def finish(wait=True):
        # built without scipy's O(nnz) validation passes (check_format / index-dtype scans): the device kernel guarantees
        # a canonical CSR (sorted, duplicate-free, no explicit zeros)
        X = PendingCSR((sm.n, sm.P), dtype=np_dtype)
        X.indptr, X.indices, X.data = (h.numpy() for h in host)
        X.has_sorted_indices = True
        X.has_canonical_format = True
        X.__dict__["_laris_wait"] = join
        if wait:
            X._laris_join()
        return X, fp
    return finish


def scores_to_csr(sm: ScoreMatrix, dtype=np.float32) -> sp.csr_matrix:
    if sm.row_nnz is None:
        raise ValueError("score matrix has no row counts")
    indptr, indices, data = K.to_host(*K.csr_compact(sm.S, sm.P, sm.row_nnz))
    return _assemble_csr(sm, indptr, indices, data, dtype)


def _assemble_csr(sm, indptr, indices, data, dtype):
    # built without scipy's O(nnz) validation passes (check_format / index-dtype scans): the device
    # kernel guarantees a canonical CSR (sorted, duplicate-free, no explicit zeros)
    X = sp.csr_matrix((sm.n, sm.P), dtype=np.dtype(dtype))
    if len(data) < 2 ** 31 - 1:
        indptr = indptr.astype(np.int32)             # scipy's own choice of index dtype for this size
    else:
        indices = indices.astype(np.int64)
    X.data, X.indices, X.indptr = data.astype(dtype, copy=False), indices, indptr
    X.has_sorted_indices = True
    X.has_canonical_format = True
    return X


def scores_from_csr(X) -> ScoreMatrix:
    indptr, indices, data = upload_csr(X)
    return ScoreMatrix(K.csr_to_dense(indptr, indices, data, X.shape[0], X.shape[1]), X.shape[1])


# --------------------------------------------------- spatial specificity ----
def _cosine(stats):
    dot, ss, tt = stats[0], stats[1], stats[2]
    den = np.sqrt(ss) * np.sqrt(tt)
    out = np.zeros_like(dot)
    np.divide(dot, den, out=out, where=den != 0)
    return out


def observed_stats(sm: ScoreMatrix, graph: SpatialGraph):
    """[5,P] fp64 host array: dot, sum S^2, sum T^2, sum S, nnz (observed graph, K4)."""
    return K.spatial_stats(sm.S, sm.P, graph.idx, graph.w, False, graph.order).cpu().numpy()


def null_graphs(graph: SpatialGraph, n_repeats, random_seed, rng="numpy"):
    """The R random graphs of the null as device tensors: (cols [R,n,k] int32, w [R,n,k] fp32, rng_state).
    'numpy' replays the reference's MT19937 stream on the host (laris/tools/_utils.py:436-446, :497-499) and uploads
    the indices; 'philox' generates them on the device."""
    n, k = graph.n, graph.k
    dev = graph.w.device
    cols = torch.empty((n_repeats, n, k), dtype=torch.int32, device=dev)
    wts = torch.empty((n_repeats, n, k), dtype=torch.float32, device=dev)
    state = None
    if rng == "numpy":
        for r, s in enumerate(_rng_compat.repeat_seeds(random_seed, n_repeats)):
            c, perm, state = _rng_compat.null_graph(n, k, s)
            cols[r].copy_(K.to_device(c.astype(np.int32), torch.int32).view(n, k))
            wts[r].copy_(K.gather_f32(graph.w.view(-1), K.to_device(perm, torch.int64)).view(n, k))
    elif rng == "philox":
        for r in range(n_repeats):
            K.null_graph_philox(n, k, graph.w, r, random_seed, out=(cols[r], wts[r]))
    else:
        raise ValueError(f"rng must be 'numpy' or 'philox', got {rng!r}")
    return cols, wts, state


def spatial_stats_all(sm: ScoreMatrix, graph: SpatialGraph, n_repeats, random_seed, rng="numpy", speculative=False):
    """Observed statistics and the R null repeats (K4a + K4b) as ONE device tensor [1+R, 5, P] fp64 - no host
    synchronisation; row 0 is the observed graph.  Returns (stats_d, rng_state)."""
    packed = packed_scores(sm, graph.idx.shape[1], 1, n_repeats, speculative)     # speculative: see packed_scores
    if packed is not None:                        # mostly-zero scores: all graphs in one pass over the packed rows
        cols = wts = state = None
        if n_repeats > 0:
            cols, wts, state = null_graphs(graph, n_repeats, random_seed, rng)
        return K.spatial_stats_sparse(sm.S, sm.P, packed[0], packed[1], graph.idx, graph.w, graph.order, cols, wts), state
    out = torch.empty((1 + n_repeats, 5, sm.P), dtype=torch.float64, device=sm.S.device)
    out[0].copy_(K.spatial_stats_batched(sm.S, sm.P, graph.idx, graph.w, graph.order)[0])
    state = None
    if n_repeats > 0:
        cols, wts, state = null_graphs(graph, n_repeats, random_seed, rng)
        out[1:].copy_(K.spatial_stats_batched(sm.S, sm.P, None, None, None, cols, wts))
    return out, state


def null_cosines(sm: ScoreMatrix, graph: SpatialGraph, n_repeats, random_seed, rng="numpy"):
    """Per-repeat cosine of the random-graph null (K4b).  Returns ([R,P], rng_state)."""
    st, state = spatial_stats_all(sm, graph, n_repeats, random_seed, rng)
    st = st.cpu().numpy()
    return np.asarray([_cosine(st[1 + r]) for r in range(n_repeats)]).reshape(n_repeats, sm.P), state


# ------------------------------------------------------------ cell types ----
def celltype_fraction(sm: ScoreMatrix, labels_d, C, n_c):
    """frac[p,c] = #{i in c: S_ip != 0} / n_c (K5b)."""
    cnt = K.celltype_nonzero_count(sm.S, sm.P, labels_d, C).cpu().numpy().astype(np.float64)
    return (cnt / n_c[:, None]).T


def cosg_regularise(cos, mu):
    """lambda*cos; lambda = cos^2/((1-mu)cos^2 + mu*sum_groups cos^2) (reference _utils.py:869-880)."""
    sq = cos * cos
    tot = sq.sum(1, keepdims=True)
    den = tot if mu == 1 else (1 - mu) * sq + mu * tot
    out = np.zeros_like(cos)
    np.divide(sq, den, out=out, where=cos != 0)
    return out * cos


def neighborhood_specificity(sm: ScoreMatrix, labels_d, C, graph: SpatialGraph, mu, ss=None, impl=None):
    """Raw P x C^2 co-localisation table (reference _utils.py:839-880), K5."""
    N = K.neighborhood_composition(labels_d, graph.idx, graph.w, C)
    num, qn2 = K.celltype_pair_contract(sm.S, sm.P, N, impl, order=graph.order)
    num, qn2 = num.cpu().numpy(), qn2.cpu().numpy()
    if ss is None:
        ss = observed_stats(sm, graph)[1]
    den = np.sqrt(ss)[:, None] * np.sqrt(qn2)[None, :]
    cos = np.zeros_like(num)
    np.divide(num, den, out=cos, where=den != 0)
    return cosg_regularise(cos, mu), cos


def cosg_gene_scores_async(xu: ExpressionU, labels_d, C, n_c, mu, expressed_pct, remove_lowly_expressed):
    """Enqueue the per (type, gene) sums of the packed expression and return ``finish()``, which downloads them and
    gives the COSG gene x group score (cosg.cosg semantics).  Host work placed between the two calls overlaps the kernel."""
    s_d, c_d, q_d = K.onehot_contract_csr(xu.indptr, xu.packed, xu.G_u, labels_d, C, row_end=xu.row_end)
    G_u = xu.G_u
    sums = K.to_host_async(s_d, c_d, q_d)            # on the host as soon as the contraction is done, whatever is queued after it

    def finish():
        s, c, q = sums()
        den = np.sqrt(q)[:, None] * np.sqrt(n_c.astype(np.float64))[None, :]
        cos = np.zeros((G_u, C))
        np.divide(s.T, den, out=cos, where=den != 0)
        score = cosg_regularise(cos, mu)
        if remove_lowly_expressed:
            score[c.T < n_c[None, :] * expressed_pct] = -1.0
        return score
    return finish


def cosg_gene_scores(xu: ExpressionU, labels_d, C, n_c, mu, expressed_pct, remove_lowly_expressed):
    """COSG gene x group score of the packed expression (cosg.cosg semantics)."""
    return cosg_gene_scores_async(xu, labels_d, C, n_c, mu, expressed_pct, remove_lowly_expressed)()


# --------------------------------------------------------------- p-values ----
def pvalue_counts(table, bg, active, n_perm, *, rng="philox", seed=0, draws=None, t0=0, only_ge=False):
    """Exceedance counts (K6).  ``table`` [groups,P] fp64 host, ``bg`` [P,nb] host."""
    t = K.to_device(np.ascontiguousarray(table, dtype=np.float64), torch.float64)
    b = K.to_device(np.ascontiguousarray(bg, dtype=np.int32), torch.int32)
    a = K.to_device(np.ascontiguousarray(active, dtype=np.uint8), torch.uint8) if active is not None else None
    d = None
    if rng == "numpy":
        if draws is None:
            raise ValueError("rng='numpy' needs host draws")
        d = K.to_device(np.ascontiguousarray(draws, dtype=np.uint8), torch.uint8)
    if n_perm == 0:
        z = np.zeros(t.shape, dtype=np.int32)
        return z, z.copy(), z.copy()
    ge, nz, both = K.pvalue_counts(t, b, n_perm, active=a, seed=seed, draws=d, t0=t0, only_ge=only_ge)
    if only_ge:
        ge = ge.cpu().numpy()
        return ge, np.zeros_like(ge), np.zeros_like(ge)
    return ge.cpu().numpy(), nz.cpu().numpy(), both.cpu().numpy()


def bh_fdr(p, sel):
    out = K.bh_fdr(K.to_device(np.ascontiguousarray(p, dtype=np.float64), torch.float64),
                   K.to_device(np.ascontiguousarray(sel, dtype=np.uint8), torch.uint8))
    return out.cpu().numpy()
