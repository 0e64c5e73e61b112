"""laris_b200.tl._utils - device-backed mirror of the reference's helper layer.

Same names, argument meaning and error behaviour as reference
``laris/tools/_utils.py``; the arithmetic runs in ``liblaris_b200.so``.  Host
code here only reshapes small per-pair / per-group tables the way the
reference's pandas code does.
"""
from __future__ import annotations

import weakref

import numpy as np
import pandas as pd
import scipy.sparse as sp
import torch

from .. import _kernels as K
from .. import _rng_compat, cosg_compat, engine
from .. import dist as _dist

# ---------------------------------------------------------------------------
# device-resident score cache: prepareLRInteraction leaves S on the GPU so that
# runLARIS does not re-upload lr_adata.X
# ---------------------------------------------------------------------------
_SCORE_CACHE: dict = {}


def _remember_scores(X, sm):
    key = id(X)
    _SCORE_CACHE[key] = (weakref.ref(X, lambda _r, key=key: _SCORE_CACHE.pop(key, None)), sm)


# the packed LR-gene expression prepareLRInteraction built stays on the device the same way (weakly keyed on the
# expression matrix object): runLARIS's gene-specificity step then needs neither a second upload of X nor a second
# selection pass.  Like the score cache it assumes the matrix object is not modified in place between the two calls.
_EXPR_CACHE: dict = {}


def _remember_expression(X, xu):
    if xu.packed is None:                  # dense-panel layout: the COSG contraction reads packed rows
        return
    key = id(X)
    _EXPR_CACHE[key] = (weakref.ref(X, lambda _r, key=key: _EXPR_CACHE.pop(key, None)), xu, X.shape, int(X.nnz))


def _expression_of(X, gene_idx):
    """(ExpressionU, packed column of every gene in ``gene_idx``) if X's LR-gene rows are resident and cover the genes."""
    hit = _EXPR_CACHE.get(id(X))
    if hit is None or hit[0]() is not X or hit[2] != X.shape or hit[3] != int(X.nnz):
        return None
    xu = hit[1]
    pos = np.searchsorted(xu.genes, gene_idx)
    pos[pos >= len(xu.genes)] = 0
    if len(xu.genes) == 0 or not np.array_equal(xu.genes[pos], gene_idx):
        return None
    return xu, (pos if xu.columns is None else np.asarray(xu.columns)[pos])


def _scores_of(lr_adata) -> engine.ScoreMatrix:
    X = lr_adata.X
    hit = _SCORE_CACHE.get(id(X))
    if hit is not None and hit[0]() is X and hit[1].S.shape[0] == X.shape[0] and hit[1].P == X.shape[1]:
        return hit[1]
    if not sp.issparse(X):
        X = sp.csr_matrix(np.asarray(X))
    sm = engine.scores_from_csr(X)
    if sp.issparse(lr_adata.X):
        _remember_scores(lr_adata.X, sm)
    return sm


_take_strings = cosg_compat.take_strings


def _log(verbose, *a):
    if verbose:
        print(*a)


def _write_obs_qc(lr_adata, sm, shard=None, percent_top=(50, 100, 200, 500)):
    """The per-cell columns scanpy.pp.calculate_qc_metrics(lr_adata, inplace=True) adds to ``obs``
    (reference laris/tools/__init__.py:493-496), computed on the device-resident scores:
    n_genes_by_counts, total_counts, their log1p, and pct_counts_in_top_N_genes for every N < n_vars
    (scanpy raises for N > n_vars; those columns are simply not written).  In a sharded run a rank only holds
    a block of pairs: counts and totals are all-reduced, the top-N percentages are omitted."""
    P = sm.P
    tops = [] if shard is not None else [N for N in percent_top if N < P]
    total, nnz, top = K.row_qc(sm.S, P, tops)
    total, nnz, top = total.cpu().numpy(), nnz.cpu().numpy(), top.cpu().numpy()
    if shard is not None:
        both = _dist.all_reduce_sum(np.stack([total, nnz.astype(np.float64)]))
        total, nnz = both[0], both[1].astype(np.int64)
    obs = lr_adata.obs
    obs["n_genes_by_counts"] = nnz
    obs["log1p_n_genes_by_counts"] = np.log1p(nnz)
    obs["total_counts"] = total.astype(np.float32)
    obs["log1p_total_counts"] = np.log1p(total).astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        for j, N in enumerate(tops):
            obs[f"pct_counts_in_top_{N}_genes"] = 100.0 * top[:, j] / total


# ---------------------------------------------------------------------------
# laris.pp re-exports (reference laris/preprocessing/__init__.py:28-38)
# ---------------------------------------------------------------------------
def _dense_dev(M):
    if sp.issparse(M):
        M = M.toarray()
    return K.to_device(np.ascontiguousarray(np.asarray(M), dtype=np.float32), torch.float32)


def _rowwise_cosine_similarity(A, B) -> np.ndarray:
    """Cosine between corresponding rows of A and B; 0 where a norm is 0.

    Reference laris/tools/_utils.py:33-101 (same ValueError on shape mismatch)."""
    if A.shape != B.shape:
        raise ValueError(f"Matrices A and B must have the same shape. Got A.shape = {A.shape}, B.shape = {B.shape}.")
    a, b = _dense_dev(A), _dense_dev(B)
    rows, cols = a.shape
    out = torch.empty(rows, dtype=torch.float64, device=a.device)
    K.check(K.lib.laris_rowwise_cosine(a.data_ptr(), b.data_ptr(), rows, cols, cols, cols, out.data_ptr(), K._stream()),
            "laris_rowwise_cosine")
    return out.cpu().numpy()


def _select_top_n(scores, n_top: int) -> np.ndarray:
    """Indices of the ``n_top`` largest scores, descending (reference :104-143;
    like the reference it raises ValueError when ``n_top > len(scores)``)."""
    scores = np.asarray(scores)
    reference_indices = np.arange(scores.shape[0], dtype=int)
    partition = np.argpartition(scores, -n_top)[-n_top:]
    return reference_indices[partition][np.argsort(scores[partition])[::-1]]


def _pairwise_row_multiply(sparse_matrix, cell_types, delimiter: str = "::"):
    """All C^2 element-wise row products, stacked; names ``a{delimiter}b``.

    Reference laris/tools/_utils.py:146-245."""
    N, M = sparse_matrix.shape
    if len(cell_types) != N:
        raise ValueError(f"Length of cell_types ({len(cell_types)}) must match number of rows in matrix ({N})")
    m = _dense_dev(sparse_matrix)
    out = torch.empty((N * N, M), dtype=torch.float32, device=m.device)
    K.check(K.lib.laris_pairwise_row_multiply(m.data_ptr(), N, M, out.data_ptr(), K._stream()),
            "laris_pairwise_row_multiply")
    names = np.array([f"{a}{delimiter}{b}" for a in cell_types for b in cell_types])
    return sp.csr_matrix(out.cpu().numpy().astype(np.float64)), names


# ---------------------------------------------------------------------------
# graphs and the random-graph null
# ---------------------------------------------------------------------------
def _coords(adata, use_rep):
    if use_rep not in adata.obsm:
        raise KeyError(f"Representation '{use_rep}' not found in adata.obsm. Available keys: {list(adata.obsm.keys())}")
    return np.ascontiguousarray(np.asarray(adata.obsm[use_rep]), dtype=np.float64)


def _build_adjacency_matrix(adata, use_rep: str = "X_spatial", n_nearest_neighbors: int = 10, sigma: float = 100.0):
    """kNN (self excluded) with w = exp(-d/sigma) as an n x n CSR (reference :343-394)."""
    return engine.build_graph(_coords(adata, use_rep), n_nearest_neighbors, False, sigma).to_scipy()


def _build_random_adjacency_matrix(adata, cellxcell, n_nearest_neighbors: int = 10, random_seed: int = 27):
    """Random columns + shuffled real weights (reference :397-446); pure index
    generation on the reference's numpy stream (also reseeds ``np.random`` like it)."""
    n, k = adata.n_obs, n_nearest_neighbors
    cols, perm, rs = _rng_compat.null_graph(n, k, random_seed)
    _rng_compat.sync_global(rs)
    return sp.csr_matrix((np.asarray(cellxcell.data)[perm], (np.repeat(np.arange(n), k), cols)), shape=cellxcell.shape)


def _graph_from_scipy(cellxcell, k):
    n = cellxcell.shape[0]
    A = sp.csr_matrix(cellxcell)
    if not (np.diff(A.indptr) == k).all():
        raise ValueError("cellxcell must have exactly n_nearest_neighbors entries per row")
    idx = K.to_device(A.indices.astype(np.int32).reshape(n, k), torch.int32)
    w64 = K.to_device(A.data.astype(np.float64).reshape(n, k), torch.float64)
    return engine.SpatialGraph(idx, w64, w64.to(torch.float32), w64, None, None, False)


def _generate_random_background(adata, cellxcell, genexcell, n_nearest_neighbors: int = 10, n_repeats: int = 30,
                                random_seed: int = 27, rng: str = "numpy"):
    """List of per-repeat cosine vectors of the random-graph null (reference :449-518)."""
    sm = engine.scores_from_csr(sp.csr_matrix(genexcell).T.tocsr())
    graph = _graph_from_scipy(cellxcell, n_nearest_neighbors)
    cos, state = engine.null_cosines(sm, graph, n_repeats, random_seed, rng)
    if state is not None:
        _rng_compat.sync_global(state)
    return [c for c in cos]


# ---------------------------------------------------------------------------
# cell-type machinery
# ---------------------------------------------------------------------------
def _group_codes(obs_col):
    """(codes int32, category names) with the reference's ordering rule (:816-836)."""
    if hasattr(obs_col, "cat"):
        cats = list(obs_col.cat.categories)
        return obs_col.cat.codes.to_numpy().astype(np.int32), cats
    uniq = obs_col.unique()
    try:
        [float(x) for x in uniq]
        cats = sorted(uniq, key=lambda x: float(x))
    except ValueError:
        cats = sorted(uniq)
    lut = {g: i for i, g in enumerate(cats)}
    return np.array([lut[g] for g in obs_col], dtype=np.int32), cats


def _compute_avg_expression(adata, groupby: str = "Leiden", genes=None, groups=None):
    """Per-category column mean as a (categories x genes) DataFrame (reference :248-340).
    Host-side: it is an O(nnz) pandas helper outside the scored hot loop."""
    if groupby not in adata.obs.columns:
        raise ValueError(f"'{groupby}' column not found in adata.obs. Please provide a valid column name.")
    codes, cats = _group_codes(adata.obs[groupby])
    X = sp.csr_matrix(adata.X)
    onehot = sp.csr_matrix((np.ones(len(codes)), (codes, np.arange(len(codes)))), shape=(len(cats), len(codes)))
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.asarray((onehot @ X).todense()) / np.bincount(codes, minlength=len(cats))[:, None]
    res = pd.DataFrame(mean, index=cats, columns=adata.var_names)
    if genes:
        missing = [g for g in genes if g not in adata.var_names]
        if missing:
            raise ValueError(f"The following genes are not found in adata.var_names: {', '.join(missing)}")
        res = res.loc[:, [g for g in adata.var_names if g in set(genes)]]
    if groups:
        missing = [g for g in groups if g not in cats]
        if missing:
            raise ValueError(f"The following groups are not found in adata.obs['{groupby}']: {', '.join(missing)}")
        res = res.loc[[g for g in cats if g in set(groups)]]
    return res


def _gene_columns(adata, names):
    idx = pd.Index(adata.var_names).get_indexer(pd.Index(names))
    if (idx < 0).any():
        missing = list(pd.Index(names)[idx < 0].unique()[:10])
        raise KeyError(f"genes not found in adata.var_names: {missing}")
    return idx.astype(np.int64)


def _calculate_ligand_receptor_specificity(adata, laris_lr, groupby, mu: float = 100, expressed_pct: float = 0.1,
                                           remove_lowly_expressed: bool = True, mask_threshold: float = 1e-6):
    """Gene x cell-type specificity of every ligand/receptor in ``laris_lr``.

    Reference :521-620 (``cosg.cosg`` -> ``indexByGene`` -> ``iqrLogNormalize``).  The
    COSG score runs on the device; the two reshaping helpers come from the real
    ``cosg`` when importable, else from laris_b200.cosg_compat (parity unpinned)."""
    lr_set = np.unique(np.hstack([laris_lr["ligand"].values, laris_lr["receptor"].values]))
    codes, cats = _group_codes(adata.obs[groupby])
    n_c = np.bincount(codes, minlength=len(cats))
    gene_idx = _gene_columns(adata, lr_set)
    hit = _expression_of(adata.X, gene_idx)
    if hit is not None:                    # rows of the LR genes are still on the device from prepareLRInteraction
        xu, rows = hit
        score = engine.cosg_gene_scores(xu, K.to_device(codes, torch.int32), len(cats), n_c, mu, expressed_pct,
                                        remove_lowly_expressed)[rows]
    else:
        xu = engine.select_genes(adata.X, gene_idx)
        score = engine.cosg_gene_scores(xu, K.to_device(codes, torch.int32), len(cats), n_c, mu, expressed_pct,
                                        remove_lowly_expressed)
    frame, order = cosg_compat.ranked_frame(lr_set, score, cats, "@@")
    table = cosg_compat.index_by_gene_of_ranking(frame, lr_set, score, [str(c) for c in cats], order, set_nan_to_zero=True,
                                                 convert_negative_one_to_zero=True)
    table = table.loc[:, [str(c) for c in cats]]
    table.columns = cats
    return cosg_compat.iqr_log_normalize(table)


def _calculate_specificity_in_spatial_neighborhood(adata, lr_adata, groupby: str, use_rep_spatial: str = "X_spatial",
                                                   number_nearest_neighbors: int = 10, mu: float = 100,
                                                   expressed_pct: float = 0.1, remove_lowly_expressed: bool = True,
                                                   return_by_group: bool = True, key_added=None,
                                                   column_delimiter: str = "@@", contract_impl=None, shard=None,
                                                   radius=None):
    """Specificity of every LR pair for every sender::receiver neighbourhood profile.

    Reference :726-977.  Stores ``lr_adata.uns[key_added]`` with the reference's
    layout (params / COSG frame / names+scores recarrays) and returns the
    normalised P x C^2 DataFrame."""
    if groupby not in adata.obs:
        raise ValueError(f"'{groupby}' not found in adata.obs")
    if use_rep_spatial not in adata.obsm:
        raise ValueError(f"'{use_rep_spatial}' not found in adata.obsm")
    codes, cats = _group_codes(adata.obs[groupby])
    C = len(cats)
    sm = _scores_of(lr_adata)
    if radius is None:
        graph = engine.build_graph(_coords(adata, use_rep_spatial), number_nearest_neighbors, False, None)
    else:
        graph = engine.build_radius_graph(_coords(adata, use_rep_spatial), radius, False, None)
    ss = sm.cache.get("ss")
    if ss is None:
        ss = engine.observed_stats(sm, graph)[1]
        sm.cache["ss"] = ss
    if shard is None:
        raw, _ = engine.neighborhood_specificity(sm, K.to_device(codes, torch.int32), C, graph, mu, ss, contract_impl)
        lr_names = np.asarray(lr_adata.var_names, dtype=object)
    else:                       # cosines of this rank's pairs, gathered; the regularisation runs over a pair's C^2 row
        _, cos = engine.neighborhood_specificity(sm, K.to_device(codes, torch.int32), C, graph, mu, ss, contract_impl)
        raw = engine.cosg_regularise(_dist.all_gather_pairs(cos, shard, axis=0), mu)
        lr_names = np.asarray(shard.names_global, dtype=object)
    pair_names = [f"{a}::{b}" for a in cats for b in cats]
    if key_added is None:
        key_added = "cosg"
    frame, order = cosg_compat.ranked_frame(lr_names, raw, pair_names, column_delimiter)
    lr_adata.uns[key_added] = {"params": dict(groupby=groupby, method="COSG", mu=mu)}
    if return_by_group:
        lr_adata.uns[key_added]["COSG"] = frame
    names_rec = np.rec.fromarrays([lr_names[order[:, j]] for j in range(C * C)],
                                  dtype=[(nm, "O") for nm in pair_names])
    scores_rec = np.rec.fromarrays([raw[order[:, j], j].astype(np.float32) for j in range(C * C)],
                                   dtype=[(nm, "float32") for nm in pair_names])
    lr_adata.uns[key_added]["names"] = names_rec
    lr_adata.uns[key_added]["scores"] = scores_rec
    table = cosg_compat.index_by_gene_of_ranking(frame, lr_names, raw, pair_names, order, column_delimiter=column_delimiter)
    return cosg_compat.iqr_log_normalize(table)


def _prepare_background_interactions(lr_adata, n_nearest_neighbors: int = 30, leaf_size: int = 40, shard=None):
    """30-NN of every LR pair in (mean, variance) space (reference :1013-1109).

    The column moments come from the device pass over S; the neighbour search
    over the P feature points is sklearn's KD-tree, exactly as in the reference
    (P is a few thousand: host-sized)."""
    from sklearn.neighbors import KDTree
    sm = _scores_of(lr_adata)
    if "sum" not in sm.cache:
        graph = engine.SpatialGraph(torch.zeros((sm.n, 1), dtype=torch.int32, device=sm.S.device), None,
                                    torch.zeros((sm.n, 1), dtype=torch.float32, device=sm.S.device), None, None, None, False)
        st = engine.observed_stats(sm, graph)
        sm.cache.update(ss=st[1], sum=st[3], nnz=st[4])
    n = sm.n
    csum, css, names = sm.cache["sum"], sm.cache["ss"], lr_adata.var_names
    if shard is not None:
        csum, css = _dist.all_gather_pairs(csum, shard), _dist.all_gather_pairs(css, shard)
        names = pd.Index(shard.names_global, dtype=object)
    mean = csum / n
    var = css / n - mean ** 2
    feats = np.column_stack([mean, var])
    kq = n_nearest_neighbors + 1
    if 2 * kq < len(feats) and kq <= 64:
        # same exact search as the spatial graphs (K1: sklearn KD-tree distances bit for bit, ties by index),
        # self included and dropped as column 0 like the reference (:1098)
        idx = K.knn_graph(K.to_device(feats, torch.float64), kq, True, want_order=False)[0].cpu().numpy().astype(np.int64)
    else:      # tiny tables (k >= P/2: sklearn switches to brute force) or more than 63 neighbours: host KD-tree
        _, idx = KDTree(feats, leaf_size=leaf_size, metric="euclidean").query(feats, k=min(kq, len(feats)))
    return pd.DataFrame(idx[:, 1:], index=names, columns=range(idx.shape[1] - 1))


def _calculate_laris_score_by_celltype(adata, lr_adata, laris_lr, groupby, use_rep_spatial: str = "X_spatial",
                                       number_nearest_neighbors: int = 10, mu: float = 100, expressed_pct: float = 0.1,
                                       remove_lowly_expressed: bool = True, mask_threshold: float = 1e-6,
                                       calculate_pvalues: bool = True, layer=None, n_nearest_neighbors: int = 30,
                                       n_permutations: int = 1000, chunk_size: int = 50000, prefilter_fdr: bool = True,
                                       prefilter_threshold: float = 0.0, score_threshold: float = 1e-6,
                                       spatial_weight: float = 1.0, use_conditional_pvalue: bool = False, *,
                                       rng: str = "numpy", random_seed: int = 27, rng_state=None, verbose: bool = True,
                                       contract_impl=None, shard=None, radius=None):
    """Sender/receiver cell-type table with permutation p-values and BH-FDR.

    Reference laris/tools/_utils.py:1113-1688; same arguments (``mask_threshold`` and
    ``layer`` are accepted and ignored there too), plus keyword-only
    ``rng`` ('numpy' replays the reference's MT19937 stream from ``rng_state``,
    'philox' uses the device generator keyed by ``random_seed``)."""
    cols6 = ["sender", "receiver", "ligand", "receptor", "interaction_name", "interaction_score"]
    codes, cats = _group_codes(adata.obs[groupby])
    if not hasattr(adata.obs[groupby], "cat"):
        raise AttributeError(f"adata.obs['{groupby}'] must be categorical (reference _compute_avg_expression)")
    C = len(cats)
    cats_arr = np.asarray(cats, dtype=object)
    n_c = np.bincount(codes, minlength=C)
    labels_d = K.to_device(codes, torch.int32)
    lr_names = np.asarray(lr_adata.var_names if shard is None else shard.names_global, dtype=object)
    P = len(lr_names)

    # ---- step 1: ligand / receptor specificity (reference :1306) ----
    _log(verbose, "\n--- Step 1: Calculating ligand and receptor cell type specificity ---")
    if len(laris_lr) == 0:
        return pd.DataFrame(columns=cols6)
    spec = _calculate_ligand_receptor_specificity(adata, laris_lr, groupby, mu, expressed_pct, remove_lowly_expressed,
                                                  mask_threshold)
    spec = spec.loc[:, cats] if list(spec.columns) != list(cats) else spec

    # ---- step 2: fraction of cells with a non-zero score (reference :1325-1337) ----
    _log(verbose, "\n--- Step 2: Calculating diffused LR-score distribution ---")
    sm = _scores_of(lr_adata)
    with np.errstate(invalid="ignore", divide="ignore"):
        frac_all = engine.celltype_fraction(sm, labels_d, C, n_c.astype(np.float64))          # [P, C]
    if shard is not None:
        frac_all = _dist.all_gather_pairs(frac_all, shard, axis=0)

    # ---- step 3: interaction scores, pair-major / receiver-major / sender fastest (:1351-1406) ----
    _log(verbose, f"\n--- Step 3: Computing interaction scores (spatial_weight={spatial_weight}) ---")
    lig = laris_lr["ligand"].to_numpy(dtype=object)
    rec = laris_lr["receptor"].to_numpy(dtype=object)
    sp_score = laris_lr["score"].to_numpy(dtype=np.float64)
    inter_names = np.array([f"{l}::{r}" for l, r in zip(lig, rec)], dtype=object)
    keep = np.isin(lig, spec.index) & np.isin(rec, spec.index)
    lig, rec, sp_score, inter_names = lig[keep], rec[keep], sp_score[keep], inter_names[keep]
    if len(lig) == 0:
        return pd.DataFrame(columns=cols6)
    pair_id = pd.Index(lr_names).get_indexer(inter_names)
    if (pair_id < 0).any():
        raise KeyError(f"interactions not found in lr_adata.var_names: {list(inter_names[pair_id < 0][:5])}")
    sl = spec.loc[lig].to_numpy(dtype=np.float64)
    sr = spec.loc[rec].to_numpy(dtype=np.float64)
    with np.errstate(invalid="ignore"):
        factor = np.ones_like(sp_score) if spatial_weight == 0 else np.power(sp_score, spatial_weight)
    frac = frac_all[pair_id]
    inter = (frac[:, :, None] * frac[:, None, :]) * (sl[:, :, None] * sr[:, None, :] * factor[:, None, None])
    npair = len(lig)
    flat = inter.transpose(0, 2, 1).reshape(-1)
    row_pair = np.repeat(np.arange(npair), C * C)
    row_recv = np.tile(np.repeat(np.arange(C), C), npair)
    row_send = np.tile(np.arange(C), npair * C)
    o = np.argsort(-flat, kind="stable")
    flat, row_pair, row_recv, row_send = flat[o], row_pair[o], row_recv[o], row_send[o]

    # ---- step 3.5: rescale (:1426-1437) ----
    top_mean = flat[: min(100, flat.size)].mean()
    if top_mean > 0:
        flat = flat * (0.1 / top_mean)

    # ---- step 4: co-localisation weight (:1448-1472) ----
    _log(verbose, "\n--- Step 4: Incorporating spatial cell type co-localization ---")
    ctc = _calculate_specificity_in_spatial_neighborhood(
        adata, lr_adata, groupby=groupby, use_rep_spatial=use_rep_spatial,
        number_nearest_neighbors=number_nearest_neighbors, mu=mu, return_by_group=True, key_added="cosg_lr",
        column_delimiter="@@", contract_impl=contract_impl, shard=shard, radius=radius)
    ctc_cols = [f"{a}::{b}" for a in cats for b in cats]
    ctc_m = ctc.loc[lr_names[pair_id], ctc_cols].to_numpy(dtype=np.float64).reshape(npair, C, C)
    flat = flat * ctc_m[row_pair, row_send, row_recv]
    o = np.argsort(-flat, kind="stable")
    flat, row_pair, row_recv, row_send = flat[o], row_pair[o], row_recv[o], row_send[o]

    # ---- step 4.5: clean-up (:1480-1489) ----
    flat = flat.astype(np.float64)
    flat[flat < score_threshold] = 0.0

    cat_taker = cosg_compat.StringTaker(cats_arr)
    res = pd.DataFrame({
        "sender": cat_taker.take(row_send), "receiver": cat_taker.take(row_recv), "interaction_score": flat,
        "ligand": _take_strings(lig, row_pair), "receptor": _take_strings(rec, row_pair),
        "interaction_name": _take_strings(inter_names, row_pair)})

    if not calculate_pvalues:
        res["p_value"] = np.nan
        res["p_value_fdr"] = np.nan
        res["nlog10_p_value_fdr"] = np.nan
        return res

    # ---- step 5.1: background sets (:1507) ----
    _log(verbose, "\n--- Step 5.1: Preparing background interaction sets ---")
    bg = _prepare_background_interactions(lr_adata, n_nearest_neighbors=n_nearest_neighbors, shard=shard).to_numpy()
    nb = bg.shape[1]

    # ---- step 5.2: permutation p-values (:1518-1596) ----
    _log(verbose, f"\n--- Step 5.2: Calculating p-values ({n_permutations:,} permutations, rng={rng}) ---")
    group_of = row_send * C + row_recv
    table = np.zeros((C * C, P))
    active = np.zeros(P, dtype=np.uint8)
    table[group_of, pair_id[row_pair]] = flat
    active[pair_id] = 1
    draws = None
    if rng == "numpy":
        rs = rng_state if rng_state is not None else np.random.RandomState(random_seed)
        draws = np.zeros((C * C, P, n_permutations), dtype=np.uint8)
        # groups in sorted (sender, receiver) NAME order, rows in table order, chunked like the reference
        gkeys = sorted({(str(cats[g // C]), str(cats[g % C]), g) for g in np.unique(group_of)})
        for _, _, g in gkeys:
            rows = np.flatnonzero(group_of == g)
            for s in range(0, len(rows), chunk_size):
                rr = rows[s:s + chunk_size]
                draws[g, pair_id[row_pair[rr]], :] = _rng_compat.pvalue_draws(rs, len(rr), n_permutations, nb)
        _rng_compat.sync_global(rs)
    elif rng != "philox":
        raise ValueError(f"rng must be 'numpy' or 'philox', got {rng!r}")
    if shard is None:
        ge, nz, both = engine.pvalue_counts(table, bg, active, n_permutations, rng=rng, seed=random_seed, draws=draws,
                                            only_ge=not use_conditional_pvalue)
    else:                       # permutations [t0, t1) on this rank; exceedance counts are integers, so the sum is exact
        pb = _dist.perm_bounds(n_permutations, shard.world)
        t0, t1 = int(pb[shard.rank]), int(pb[shard.rank + 1])
        part = engine.pvalue_counts(table, bg, active, t1 - t0, rng=rng, seed=random_seed, t0=t0,
                                    draws=None if draws is None else np.ascontiguousarray(draws[:, :, t0:t1]),
                                    only_ge=not use_conditional_pvalue)
        ge, nz, both = _dist.all_reduce_sum(np.stack(part).astype(np.int64))
    if use_conditional_pvalue:
        p_tab = np.ones_like(table)
        pos = table > 0
        p_tab[pos] = (both[pos] + 1) / (nz[pos] + 1)
    else:
        p_tab = (ge + 1) / (n_permutations + 1)
    res["p_value"] = p_tab[group_of, pair_id[row_pair]]

    # ---- step 5.3: BH-FDR per group (:1617-1658) ----
    _log(verbose, "\n--- Step 5.3: Applying FDR correction (Benjamini-Hochberg) ---")
    sel = np.broadcast_to(active[None, :].astype(bool), table.shape).copy()
    if prefilter_fdr:
        sel &= table > prefilter_threshold
    fdr_tab = engine.bh_fdr(np.round(p_tab, 6), sel.astype(np.uint8))
    fdr = np.nan_to_num(np.minimum(fdr_tab[group_of, pair_id[row_pair]], 1.0), nan=1.0)
    res["p_value_fdr"] = fdr
    res["nlog10_p_value_fdr"] = np.maximum(-np.log10(fdr + 1e-10), 0.0)
    return res


__all__ = [
    "_rowwise_cosine_similarity",
    "_select_top_n",
    "_calculate_laris_score_by_celltype",
    "_calculate_specificity_in_spatial_neighborhood",
]
