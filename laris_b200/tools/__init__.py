"""laris_b200.tl - drop-in for the reference's ``laris.tl`` scoring path.

``prepareLRInteraction`` and ``runLARIS`` keep the reference's signatures,
defaults, error behaviour and AnnData layouts
(reference laris/tools/__init__.py:29-119 and :121-593); the arithmetic runs in
hand-written sm_100a kernels behind the C ABI of include/laris_b200.h.
Extra keyword-only arguments (``rng``, ``verbose``, ``dtype``, ``contract_impl``)
default to the reference's behaviour.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse as sp
import torch

from .. import _kernels as K
from .. import _rng_compat, engine
from .. import dist as _dist
from .._anndata import make_anndata
from . import _utils
from ._utils import _log


def _lr_units(adata, lr_df, complex_sep):
    """Like ``_lr_gene_indices`` for tables whose names may be multi-subunit complexes ``A{sep}B{sep}...``:
    returns (ligand ids, receptor ids, complexes) where an id >= n_genes names ``complexes[id - n_genes]``."""
    G = adata.shape[1]
    lig_n, rec_n = lr_df["ligand"].astype(str).to_numpy(), lr_df["receptor"].astype(str).to_numpy()
    names = pd.unique(np.concatenate([lig_n, rec_n]))
    parts = {nm: [s for s in nm.split(complex_sep) if s] for nm in names}
    cx_names = [nm for nm in names if len(parts[nm]) > 1]           # "A+" is just gene A
    cx_id = {nm: G + c for c, nm in enumerate(cx_names)}
    complexes = [_utils._gene_columns(adata, parts[nm]) for nm in cx_names]
    plain = [nm for nm in names if len(parts[nm]) <= 1]
    simple = {nm: g for nm, g in zip(plain, _utils._gene_columns(adata, [(parts[nm] or [nm])[0] for nm in plain]))}
    ids = {**simple, **cx_id}
    return (np.array([ids[nm] for nm in lig_n], dtype=np.int64), np.array([ids[nm] for nm in rec_n], dtype=np.int64),
            complexes)


def _lr_gene_indices(adata, lr_df):
    """Gene columns of every ligand / receptor.  The reference's
    argsort/searchsorted lookup (laris/tools/__init__.py:95-97) silently maps a
    missing name to its lexicographic successor; here a missing gene raises."""
    return _utils._gene_columns(adata, lr_df["ligand"].astype(str)), _utils._gene_columns(adata, lr_df["receptor"].astype(str))


def prepareLRInteraction(adata, lr_df, number_nearest_neighbors: int = 10, use_rep_spatial: str = "X_spatial", *,
                         dtype=np.float32, layout: str = "auto", radius=None, complex_sep=None,
                         complex_rule: str = "min", csr_dev=None, to_host: bool = True):
    """Per-cell diffused ligand-receptor scores as a new AnnData (cells x LR pairs).

    Reference laris/tools/__init__.py:29-119: kNN graph incl. self with
    w = 1/exp(d/(mean(d)/2)); D = W X; S[i,p] = D[i,l_p] D[i,r_p] masked to cells
    expressing the ligand or the receptor.  Returns a new AnnData with ``.X`` the
    sparse n x P score matrix (exact zeros pruned), ``obs``/``obsm`` copied and
    ``var_names`` = "ligand::receptor".  The dense scores stay resident on the GPU
    for a following ``runLARIS`` call.  ``layout`` picks the device layout of the LR-gene
    expression ('dense' rows, packed 'sparse' rows, or 'auto' by the fraction of non-zeros);
    both give the same result.  ``radius`` (extension, not in the reference): use the radius graph
    {j: |x_i - x_j| <= radius}, self included, instead of the kNN graph; weights follow the same
    1/exp(d/(mean(d)/2)) rule over the stored distances.  ``complex_sep`` (extension): a ligand/receptor
    name containing this separator (e.g. ``"TGFBR1+TGFBR2"`` with ``complex_sep="+"``) is a multi-subunit
    complex; its diffused expression is the ``complex_rule`` ('min' | 'gmean') of its subunits' and it counts as
    expressed in a cell only if every subunit is.  Without it, complexes are pre-expanded rows as in the
    reference's tutorial.  (The spatial-specificity step of ``runLARIS`` takes such rows as they are; its
    cell-type step looks ligand / receptor names up as genes, like the reference, so it needs expanded rows.)
    ``csr_dev`` (extension): ``adata.X`` already on the device as ``(indptr int64, indices int32, data fp32)``
    tensors (e.g. from ``laris_b200.dist.upload_csr_replicated``) - nothing is uploaded.  ``to_host=False``
    (extension): skip the CSR compaction and download; ``.X`` of the returned (AnnDataLite) object is then a
    ``ResidentScores`` handle that ``runLARIS`` consumes directly and ``.X.tocsr()`` materialises on demand.

    With ``to_host=True`` (default) ``.X`` is a ``scipy.sparse.csr_matrix`` (subclass ``engine.PendingCSR``) that is
    returned while its arrays are still being copied from the device: the first access to them waits for the copy.
    Limits of the compiled kernels (a ``LarisError`` names the limit when one is exceeded): at most 64 neighbours per
    cell (``number_nearest_neighbors`` or the maximum degree of a radius graph) and at most 65533 distinct LR-gene /
    complex columns.
    """
    xy = _utils._coords(adata, use_rep_spatial)
    # schedule: the expression matrix goes up on the copy stream while the host resolves the LR table and the graph is
    # built; the result comes down while the host prepares obs / var / obsm of the new AnnData
    if csr_dev is None:
        csr_dev, join_upload = engine.upload_csr_async(adata.X)
    else:
        K.require_cuda()
        join_upload = lambda: csr_dev                                      # noqa: E731
    if radius is None:                                  # the device builds the graph while the host resolves the LR table
        graph = engine.build_graph(xy, number_nearest_neighbors, True, None)
    else:
        graph = engine.build_radius_graph(xy, radius, True, None)
    complexes = None
    if complex_sep:
        lig_g, rec_g, complexes = _lr_units(adata, lr_df, complex_sep)
    else:
        lig_g, rec_g = _lr_gene_indices(adata, lr_df)
    csr_dev = join_upload()
    expr_fp = engine.device_fingerprint_async(csr_dev[2])      # read lazily, at the first cache lookup
    xu, lig_c, rec_c = engine.select_lr_genes(adata.X, lig_g, rec_g, csr_dev=csr_dev, layout=layout,
                                              complexes=complexes, rule=complex_rule)
    sm = engine.lr_scores(xu, graph, lig_c, rec_c)
    finish = engine.scores_to_csr_async(sm, dtype) if to_host else None
    lr_names = (lr_df["ligand"].astype(str) + "::" + lr_df["receptor"].astype(str)).to_numpy(dtype=object)
    obs, var = adata.obs.copy(), pd.DataFrame(index=pd.Index(lr_names, dtype=object))
    # device-resident coordinates are shared, not cloned (nothing on this path writes to them)
    obsm = {k: (v if torch.is_tensor(v) else np.array(v, copy=True)) for k, v in adata.obsm.items()}
    if to_host:
        # the matrix is handed over while its arrays are still coming down (engine.PendingCSR): whoever reads them first
        # waits for the copies; a following runLARIS does not, it scores the device-resident copy
        X, score_fp = finish(wait=False)
        lr_adata = make_anndata(X, obs=obs, var=var, obsm=obsm)
        _utils._remember_scores(lr_adata.X, sm, score_fp)
    else:
        from .._anndata import AnnDataLite
        lr_adata = AnnDataLite(_utils.ResidentScores(sm), obs, var, obsm)
    _utils._remember_expression(adata.X, xu, expr_fp)
    return lr_adata


def runLARIS(lr_adata, adata=None, use_rep: str = "X_spatial", n_nearest_neighbors: int = 10, random_seed: int = 27,
             n_repeats: int = 3, mu: float = 1, sigma: float = 100, remove_lowly_expressed: bool = True,
             expressed_pct: float = 0.1, n_cells_expressed_threshold: int = 100, n_top_lr: int = 4000,
             by_celltype: bool = True, groupby: str = "CellTypes", use_rep_spatial: str = "X_spatial",
             number_nearest_neighbors: int = 10, mu_celltype: float = 100, expressed_pct_celltype: float = 0.1,
             remove_lowly_expressed_celltype: bool = True, mask_threshold: float = 1e-6, calculate_pvalues: bool = True,
             layer_celltype=None, n_neighbors_permutation: int = 30, n_permutations: int = 1000, chunk_size: int = 50000,
             prefilter_fdr: bool = True, prefilter_threshold: float = 0.0, score_threshold: float = 1e-6,
             spatial_weight: float = 1.0, use_conditional_pvalue: bool = False, *, rng: str = "numpy",
             verbose: bool = True, contract_impl=None, radius=None, output: str = "frames", label_permutations: int = 0):
    """Spatially-specific LR pairs and (optionally) the cell-type interaction table.

    Reference laris/tools/__init__.py:121-593, same arguments and returns
    (``laris_lr`` or ``(laris_lr, celltype_results)``), same in-place writes to
    ``lr_adata.var`` / ``lr_adata.uns['cosg_lr']``.  ``n_top_lr`` is clamped to the
    number of LR pairs (the reference raises for P < 4000);
    ``remove_lowly_expressed``/``expressed_pct`` are accepted and unused, as in the
    reference body.  ``rng='numpy'`` replays the reference's random stream
    bit-exactly; ``rng='philox'`` generates the nulls on the device.  ``radius`` (extension): both spatial graphs
    (step 1 with ``exp(-d/sigma)`` weights, the neighbourhood profile with ``exp(-d/(mean(d)/2))``) are radius
    graphs without self edges instead of kNN graphs; the random-graph null then re-deals the rows' padded
    fixed-degree weight slots (zero padding included), which keeps the reference's construction for a graph whose
    degree varies.  ``output='arrays'`` (extension): skip the materialisation of the results - the per-cell QC
    columns, ``uns['cosg_lr']`` and the sorted long DataFrame - and return the cell-type step as a dict of device
    tensors ([pair, sender*C + receiver] layout); the ranking DataFrame and the ``var`` columns are still written.
    ``label_permutations=T`` (extension, default 0 = the reference's behaviour): additionally relabel the cells T times
    and add ``p_value_label_null`` - the label-permutation p-value of the row's co-localisation cosine
    (``labelPermutationNull``).

    Limits of the compiled kernels (a ``LarisError`` / ``ValueError`` names the limit when one is exceeded):
    ``n_nearest_neighbors`` / ``number_nearest_neighbors`` (or the maximum radius-graph degree) <= 64;
    ``n_neighbors_permutation`` <= 256; at most 16384 LR pairs per sender-receiver group in the BH-FDR step; the default
    cell-type contraction takes <= 64 cell types (more fall back to the fp32 tile kernel automatically).
    """
    if by_celltype and adata is None:
        raise ValueError("Parameter 'adata' must be provided when by_celltype=True. "
                         "adata should contain gene expression and cell type annotations.")
    if use_rep not in lr_adata.obsm:
        raise KeyError(f"Representation '{use_rep}' not found in lr_adata.obsm. "
                       f"Available keys: {list(lr_adata.obsm.keys())}")
    if output not in ("frames", "arrays"):
        raise ValueError(f"output must be 'frames' or 'arrays', got {output!r}")
    _log(verbose, "\n" + "=" * 70 + "\nLARIS ANALYSIS\n" + "=" * 70)
    _log(verbose, f"\nInput data: {lr_adata.shape[0]} cells x {lr_adata.shape[1]} LR pairs")

    # ---- step 1: spatial specificity (reference :459-490) ----
    _log(verbose, "\n--- Step 1: Calculating spatial specificity scores ---")
    # multi-GPU: this rank holds a block of LR pairs (laris_b200.dist.prepare_sharded); per-pair results are
    # all-gathered in global pair order, so everything from the ranking on is identical on every rank
    shard = _dist.shard_of(lr_adata)
    own = slice(None) if shard is None else slice(shard.start, shard.stop)
    sm = _utils._scores_of(lr_adata)
    if radius is None:
        graph = engine.build_graph(_utils._coords(lr_adata, use_rep), n_nearest_neighbors, False, sigma)
    else:
        graph = engine.build_radius_graph(_utils._coords(lr_adata, use_rep), radius, False, sigma)
    # observed statistics + all null repeats: one device tensor [1+R, 5, P_local], one gather, one download
    # (speculative sizing of the packed rows only without sharding: a redo on some ranks would unbalance the collectives)
    st_d, rng_state = engine.spatial_stats_all(sm, graph, n_repeats, random_seed, rng, speculative=shard is None)
    sm.cache["stats_d"] = st_d[0]
    sm.cache["graph"] = graph                 # the cell-type step re-uses the neighbour search when it can
    # the statistics come down through a copy queued NOW (K.to_host_async), so that waiting for them below does not also
    # wait for the cell-type work queued in between
    if shard is not None:
        stats_get = K.to_host_async(_dist.all_gather_pairs_dev(st_d, shard, axis=-1), st_d[0].contiguous())
    else:
        stats_get = K.to_host_async(st_d)
    pre = None
    if by_celltype:
        # queue the ranking-independent part of the cell-type step BEFORE waiting for the statistics: the device then has
        # work while the host ranks the pairs (reference order of operations is unchanged where it matters: rng='numpy'
        # consumes the random stream in the same order, nothing here draws from it)
        if groupby not in adata.obs:
            raise ValueError(f"Cell type column '{groupby}' not found in adata.obs. "
                             f"Available columns: {list(adata.obs.columns)}")
        pre = _utils._celltype_prefetch(adata, lr_adata, groupby, use_rep_spatial, number_nearest_neighbors, mu_celltype,
                                        expressed_pct_celltype, remove_lowly_expressed_celltype, contract_impl, shard, radius)
    if not engine.packed_ok(sm):
        # the packed-row pass ran on an entry array sized from the previous matrix of this shape and this one is denser:
        # its statistics are incomplete - redo them with the exact size (the cell-type work queued above is unaffected)
        # (the random graphs are regenerated from random_seed: same graphs, same continued rng state)
        st_d, rng_state = engine.spatial_stats_all(sm, graph, n_repeats, random_seed, rng)
        sm.cache["stats_d"] = st_d[0]
        if pre is not None and pre.get("cnt_from_packed"):             # K5b counted the truncated rows too
            pre["cnt_d"] = K.celltype_nonzero_count(sm.S, sm.P, pre["labels"][3], len(pre["labels"][1]))
        if shard is not None:
            stats_get = K.to_host_async(_dist.all_gather_pairs_dev(st_d, shard, axis=-1), st_d[0].contiguous())
        else:
            stats_get = K.to_host_async(st_d)
    if shard is not None:
        st_all, st_local = stats_get()
        st = st_all[0]
        sm.cache.update(sum_global=st[3], ss_global=st[1])
    else:
        st_all = stats_get()
        st = st_local = st_all[0]
    sm.cache.update(ss=st_local[1], sum=st_local[3], nnz=st_local[4])
    rand_each = np.asarray([engine._cosine(st_all[1 + r]) for r in range(n_repeats)]).reshape(n_repeats, st.shape[1])
    gsp = engine._cosine(st)
    random_gsp = np.mean(rand_each, axis=0)
    gsp_score = gsp - mu * random_gsp
    # the reference's three var columns (:488-490), the QC column the ranking consumes (:493-504 via scanpy) and its cheap
    # siblings, written in one go
    n = lr_adata.shape[0]
    nnz_own, sum_own = st[4][own], st[3][own]
    lr_adata.var = lr_adata.var.assign(
        LRSS_Target=gsp[own], LRSS_Random=random_gsp[own], LR_SpatialSpecificity=gsp_score[own],
        n_cells_by_counts=nnz_own.astype(np.int64), mean_counts=sum_own / n, pct_dropout_by_counts=100.0 * (1.0 - nnz_own / n),
        total_counts=sum_own, log1p_mean_counts=np.log1p(sum_own / n), log1p_total_counts=np.log1p(sum_own))
    # the per-cell QC columns: kernel and download queued here, written to obs just before returning (the host does not
    # wait for the device at this point)
    qc_finish = _utils._write_obs_qc_async(lr_adata, sm, shard) if output == "frames" else (lambda: None)

    if pre is not None and calculate_pvalues:
        # the background sets only need the column moments that just arrived: queue their search now, K6 finds them
        # on the device
        pre["bg"] = (n_neighbors_permutation, _utils._background_sets_device(lr_adata, n_neighbors_permutation, shard=shard))

    # ---- ranking (reference :499-526): sort by score (descending), penalise pairs seen in too few cells, keep the top ----
    names_all = np.asarray(lr_adata.var_names if shard is None else shard.names_global, dtype=object)
    by_score = np.argsort(-gsp_score, kind="stable")                 # = sort_values(ascending=False); NaN scores last
    ranking = gsp_score[by_score].copy()
    ranking[st[4][by_score].astype(np.int64) < n_cells_expressed_threshold] = np.min(ranking) - 0.001
    top = _utils._select_top_n(ranking, min(int(n_top_lr), len(ranking)))
    names = names_all[by_score][top]
    ligs = [nm.split("::")[0] for nm in names]
    recs = [nm.split("::")[1] for nm in names]
    laris_lr = pd.DataFrame({"ligand": ligs, "receptor": recs, "score": ranking[top]})
    laris_lr.index = pd.Index([f"{l}::{r}" for l, r in zip(ligs, recs)], dtype=object)
    if pre is not None:
        pre["pair_id"] = np.asarray(by_score[top], dtype=np.int64)       # rows of laris_lr as positions in names_all
    laris_lr["Rank"] = np.arange(len(laris_lr))
    _log(verbose, f"  Identified {len(laris_lr)} top spatially-specific LR pairs")

    if not by_celltype:
        if rng_state is not None:
            _rng_compat.sync_global(rng_state)
        qc_finish()
        return laris_lr

    # ---- step 2: cell-type table (reference :536-572) ----
    if groupby not in adata.obs:
        raise ValueError(f"Cell type column '{groupby}' not found in adata.obs. "
                         f"Available columns: {list(adata.obs.columns)}")
    celltype_results = _utils._calculate_laris_score_by_celltype(
        adata=adata, lr_adata=lr_adata, laris_lr=laris_lr, groupby=groupby, use_rep_spatial=use_rep_spatial,
        number_nearest_neighbors=number_nearest_neighbors, mu=mu_celltype, expressed_pct=expressed_pct_celltype,
        remove_lowly_expressed=remove_lowly_expressed_celltype, mask_threshold=mask_threshold,
        calculate_pvalues=calculate_pvalues, layer=layer_celltype, n_nearest_neighbors=n_neighbors_permutation,
        n_permutations=n_permutations, chunk_size=chunk_size, prefilter_fdr=prefilter_fdr,
        prefilter_threshold=prefilter_threshold, score_threshold=score_threshold, spatial_weight=spatial_weight,
        use_conditional_pvalue=use_conditional_pvalue, rng=rng, random_seed=random_seed, rng_state=rng_state,
        verbose=verbose, contract_impl=contract_impl, shard=shard, radius=radius, output=output,
        label_permutations=label_permutations, _pre=pre)
    qc_finish()
    if output == "frames":
        _log(verbose, f"\nResults: {len(laris_lr)} LR pairs, {len(celltype_results):,} cell type combinations")
    return laris_lr, celltype_results


def labelPermutationNull(lr_adata, adata, groupby: str = "CellTypes", use_rep_spatial: str = "X_spatial",
                         number_nearest_neighbors: int = 10, n_permutations: int = 100, random_seed: int = 27, *,
                         contract_impl=None, radius=None):
    """EXTENSION (not in the reference): label-permutation p-values of the sender::receiver co-localisation cosine of
    every LR pair, as a P x C^2 DataFrame (also stored in ``lr_adata.uns['laris_label_null']``).
    See ``laris_b200.tl._utils._label_permutation_null``."""
    table = _utils._label_permutation_null(adata, lr_adata, groupby, use_rep_spatial, number_nearest_neighbors,
                                           n_permutations, random_seed, contract_impl=contract_impl,
                                           shard=_dist.shard_of(lr_adata), radius=radius)
    lr_adata.uns["laris_label_null"] = {"params": dict(groupby=groupby, n_permutations=int(n_permutations),
                                                       random_seed=int(random_seed)), "p_value": table}
    return table


__all__ = ["prepareLRInteraction", "runLARIS"]
