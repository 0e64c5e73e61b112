"""numpy/scipy/sklearn restatement of the reference's LR scoring path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference lines it follows (paths relative to the reference repo root).  The
reference delegates its arithmetic to scikit-learn / scipy / numpy, which are
present in this image, so those same third-party routines are called here the
way the reference calls them; the glue between them is restated.

Pinning: ``oracle/make_golden.py`` runs the unmodified reference (loaded by
``oracle.ref_loader``) and this file on the same seeded inputs in the build
container, asserts equality, and writes ``tests/golden/*.npz``; the CPU test
suite re-checks this file against those fixtures wherever it runs.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import rng as orng

# ---------------------------------------------------------------------------
# a1-a3  spatial kNN graph + distance-decay weights
# ---------------------------------------------------------------------------


def knn_graph(xy, k, include_self):
    """Exact Euclidean kNN, rows ascending by distance.

    laris/tools/__init__.py:83-88 (include_self=True) and
    laris/tools/_utils.py:385-390, :847-852 (include_self=False) call
    ``sklearn.neighbors.kneighbors_graph(mode='distance')``; the same sklearn
    KD-tree query is used here.  Returns ``(indices int64 [n,k], dist f64 [n,k])``
    which is the CSR's ``indices``/``data`` reshaped (indptr = arange(0,nk+1,k)).
    """
    from sklearn.neighbors import NearestNeighbors
    xy = np.asarray(xy, dtype=np.float64)
    nn = NearestNeighbors(n_neighbors=k).fit(xy)
    if include_self:
        dist, ind = nn.kneighbors(xy, n_neighbors=k)
    else:
        dist, ind = nn.kneighbors(None, n_neighbors=k)   # sklearn queries k+1 and drops self
    return ind.astype(np.int64), dist.astype(np.float64)


def knn_bruteforce(xy, k, include_self, block=2048):
    """Arithmetic restatement of the KD-tree result, for pinning tie rules.

    d^2 = ((0 + dx*dx) + dy*dy [+ dz*dz]) in float64 without FMA, then sqrt
    (sklearn/metrics/_dist_metrics euclidean rdist).  Ordering by
    (d^2, index).  exclude-self follows sklearn/neighbors/_base.py kneighbors:
    query k+1, drop the entry equal to the row id, or the first entry if the
    row id is absent.
    """
    xy = np.asarray(xy, dtype=np.float64)
    n, d = xy.shape
    kk = k if include_self else k + 1
    ind = np.empty((n, kk), np.int64)
    d2o = np.empty((n, kk), np.float64)
    for s in range(0, n, block):
        q = xy[s:s + block]
        d2 = np.zeros((q.shape[0], n))
        for a in range(d):
            t = q[:, a, None] - xy[None, :, a]
            d2 = d2 + t * t
        order = np.lexsort((np.broadcast_to(np.arange(n), d2.shape), d2), axis=1)[:, :kk]
        ind[s:s + block] = order
        d2o[s:s + block] = np.take_along_axis(d2, order, 1)
    dist = np.sqrt(d2o)
    if include_self:
        return ind, dist
    rows = np.arange(n)[:, None]
    keep = ind != rows
    missing = keep.all(1)
    keep[missing, 0] = False
    return ind[keep].reshape(n, k), dist[keep].reshape(n, k)


def radius_graph(xy, radius, include_self):
    """EXTENSION oracle (north-star item 1; the reference has no radius graph, SURVEY.md "gaps"):
    ``sklearn.neighbors.radius_neighbors_graph(mode='distance')`` as canonical CSR (column-sorted, explicit
    zero distance on the diagonal when ``include_self``).  Returns ``(indptr, indices, dist)``."""
    from sklearn.neighbors import radius_neighbors_graph
    A = sp.csr_matrix(radius_neighbors_graph(np.asarray(xy, dtype=np.float64), radius, mode="distance",
                                             include_self=bool(include_self)))
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)


def lr_scores_csr_graph(X, W, lig_idx, rec_idx):
    """``lr_scores`` for an arbitrary n x n CSR weight matrix W (same arithmetic, laris/tools/__init__.py:91-117)."""
    D = (W @ sp.csr_matrix(X)).tocsc()
    Xc = sp.csc_matrix(X)
    mask = (Xc[:, lig_idx] != 0).maximum(Xc[:, rec_idx] != 0)
    S = D[:, lig_idx].multiply(D[:, rec_idx]).multiply(mask).tocsr()
    S.eliminate_zeros()
    S.sort_indices()
    return S


def lr_scores_complex(X, W, lig_sets, rec_sets, rule="min"):
    """EXTENSION oracle (north-star item 3; no reference behaviour, SURVEY.md "gaps"): LR scores where a ligand /
    receptor may be a multi-subunit complex (a list of gene columns).  D = W X as in the reference
    (laris/tools/__init__.py:91-92); a unit's diffused expression is the min ('min') or the geometric mean of
    max(v,0) ('gmean') of its subunits' D; a unit is expressed in a cell iff every subunit has X != 0; the score is
    D_L * D_R masked to cells expressing the ligand unit or the receptor unit (:100, :110-117 for single genes)."""
    Xd = np.asarray(sp.csr_matrix(X).todense(), dtype=np.float64)
    D = np.asarray((sp.csr_matrix(W) @ sp.csr_matrix(X)).todense(), dtype=np.float64)

    def unit(cols):
        v = D[:, cols]
        if len(cols) == 1:                 # a single gene is never aggregated: the reference's own arithmetic
            a = v[:, 0]
        elif rule == "min":
            a = v.min(1)
        else:
            v = np.maximum(v, 0.0)
            with np.errstate(divide="ignore"):
                a = np.where((v > 0).all(1), np.exp(np.log(np.where(v > 0, v, 1.0)).mean(1)), 0.0)
        return a, (Xd[:, cols] != 0).all(1)

    out = np.zeros((X.shape[0], len(lig_sets)))
    for p, (ls, rs) in enumerate(zip(lig_sets, rec_sets)):
        (a, fa), (b, fb) = unit(list(ls)), unit(list(rs))
        out[:, p] = a * b * (fa | fb)
    S = sp.csr_matrix(out)
    S.eliminate_zeros()
    S.sort_indices()
    return S


def decay_weights(dist, sigma=None):
    """w = 1/exp(d/s); s = mean(all stored d)/2 or the constant ``sigma``.

    laris/tools/__init__.py:89, laris/tools/_utils.py:392, :853.
    """
    s = (np.mean(dist) / 2.0) if sigma is None else sigma
    return 1.0 / np.exp(dist / s)


def graph_csr(ind, w, n):
    k = ind.shape[1]
    return sp.csr_matrix((w.ravel(), ind.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))


# ---------------------------------------------------------------------------
# a4-a5  diffusion + per-cell LR score
# ---------------------------------------------------------------------------


def gene_index(var_names, names):
    """argsort/searchsorted lookup, laris/tools/__init__.py:95-97."""
    var_names = np.asarray(var_names)
    sorter = np.argsort(var_names)
    return sorter[np.searchsorted(var_names, np.asarray(names), sorter=sorter)]


def lr_scores(X, ind, w, lig_idx, rec_idx):
    """S[i,p] = D[i,l_p] * D[i,r_p] * [X[i,l_p] != 0  or  X[i,r_p] != 0],  D = W X.

    laris/tools/__init__.py:91-92 (diffusion over all genes), :100 (gather +
    product), :110-117 (expression mask).  Returns CSR n x P (float64) with
    exact zeros pruned, like the reference's sparse products.
    """
    n = X.shape[0]
    W = graph_csr(ind, w, n)
    D = (W @ sp.csr_matrix(X)).tocsc()
    Dl, Dr = D[:, lig_idx], D[:, rec_idx]
    Xc = sp.csc_matrix(X)
    mask = (Xc[:, lig_idx] != 0).maximum(Xc[:, rec_idx] != 0)
    S = Dl.multiply(Dr).multiply(mask).tocsr()
    S.eliminate_zeros()
    S.sort_indices()
    return S


# ---------------------------------------------------------------------------
# a6-a8  spatial specificity + random-graph null + ranking
# ---------------------------------------------------------------------------


def column_cosine(S, T):
    """Per-column <S,T>/(|S||T|), 0 where the denominator is 0.

    laris/tools/_utils.py:81-101 (applied there to rows of the transposes).
    """
    S, T = sp.csc_matrix(S), sp.csc_matrix(T)
    dot = np.asarray(S.multiply(T).sum(0)).ravel()
    den = np.sqrt(np.asarray(S.multiply(S).sum(0)).ravel()) * \
        np.sqrt(np.asarray(T.multiply(T).sum(0)).ravel())
    out = np.zeros_like(dot)
    np.divide(dot, den, out=out, where=den != 0)
    return out


def observed_cosine(S, ind, w):
    """gsp_p = cos(S[:,p], (W2 S)[:,p]) with W2 NOT normalised;
    laris/tools/__init__.py:468-470."""
    W = graph_csr(ind, w, S.shape[0])
    return column_cosine(S, W @ sp.csr_matrix(S))


def null_graph_csr(n, k, cols, wperm):
    """Random graph: COO->CSR sums duplicate (row,col), then L1 row-normalise.

    laris/tools/_utils.py:436-446 (construction), :512 (normalize l1).
    """
    from sklearn.preprocessing import normalize
    rows = np.repeat(np.arange(n), k)
    A = sp.csr_matrix((wperm, (rows, cols)), shape=(n, n))
    return normalize(A, axis=1, norm="l1")


def null_cosine(S, k, cols, wperm):
    """cos(S[:,p], (W_rand S)[:,p]) for one repeat; laris/tools/_utils.py:514-516."""
    A = null_graph_csr(S.shape[0], k, cols, wperm)
    return column_cosine(S, A @ sp.csr_matrix(S))


def null_graphs(n, k, w, n_repeats, random_seed, mode):
    """Yield (cols, wperm) per repeat for ``mode`` in {'numpy','philox'}.

    numpy : laris/tools/_utils.py:496-499 + :436-442 replayed on a private
            RandomState (the reference uses the global one).
    philox: the device generator's CPU twin (oracle/rng.py).
    """
    wflat = np.asarray(w, dtype=np.float64).ravel()
    if mode == "numpy":
        for s in orng.numpy_repeat_seeds(random_seed, n_repeats):
            cols, perm, _ = orng.numpy_null_graph(n, k, s)
            yield cols, wflat[perm]
    elif mode == "philox":
        for r in range(n_repeats):
            cols = orng.philox_null_cols(n, n * k, r, random_seed)
            perm = orng.feistel_permutation(n * k, r, random_seed)
            yield cols, wflat[perm]
    else:
        raise ValueError(mode)


def spatial_specificity(S, ind, w, n_repeats, random_seed, mu, mode="numpy"):
    """(gsp, rand, score): laris/tools/__init__.py:468-485."""
    n, k = ind.shape
    gsp = observed_cosine(S, ind, w)
    reps = [null_cosine(S, k, c, wp) for c, wp in null_graphs(n, k, w, n_repeats, random_seed, mode)]
    rand = np.mean(reps, axis=0)
    return gsp, rand, gsp - mu * rand


def select_top_n(scores, n_top):
    """argpartition + argsort[::-1]; laris/tools/_utils.py:139-143."""
    scores = np.asarray(scores)
    part = np.argpartition(scores, -n_top)[-n_top:]
    return part[np.argsort(scores[part])[::-1]]


def rank_pairs(score, nnz_per_pair, n_cells_expressed_threshold, n_top_lr):
    """Sort desc, penalise low-count pairs, keep the top; laris/tools/__init__.py:499-517.

    Returns (pair indices in rank order, their ranking scores).
    """
    order = np.argsort(-score, kind="stable")           # reference: pandas sort_values (quicksort)
    ranked = score[order].copy()
    ranked[nnz_per_pair[order] < n_cells_expressed_threshold] = ranked.min() - 0.001
    top = select_top_n(ranked, min(n_top_lr, ranked.size))
    return order[top], ranked[top]


# ---------------------------------------------------------------------------
# a9-a10  cell-type contractions
# ---------------------------------------------------------------------------


def celltype_fraction(S, labels, C):
    """frac[p,c] = #{i in c : S_ip != 0} / n_c; laris/tools/_utils.py:1325-1337, :334-338."""
    B = sp.csr_matrix(S, copy=True)
    B.eliminate_zeros()
    B.data = np.ones_like(B.data)
    onehot = sp.csr_matrix((np.ones(len(labels)), (labels, np.arange(len(labels)))), shape=(C, len(labels)))
    counts = np.asarray((onehot @ B).todense())
    n_c = np.bincount(labels, minlength=C).astype(np.float64)
    return (counts / n_c[:, None]).T


def neighborhood_composition(labels, ind, w, C):
    """N[a,i] = sum of w_ij over neighbours j of i with type a; laris/tools/_utils.py:839-856."""
    n = len(labels)
    onehot = sp.csr_matrix((np.ones(n), (labels, np.arange(n))), shape=(C, n))
    return np.asarray((onehot @ graph_csr(ind, w, n).T).todense())


def celltype_pair_cosine(S, N):
    """cos[p,(a,b)] between S[:,p] and Q_(ab) = N[a,:]*N[b,:]; a-major pair order.

    laris/tools/_utils.py:859-866 (+ :208-245 for the row products).
    """
    from sklearn.metrics.pairwise import cosine_similarity
    C = N.shape[0]
    Q = (N[:, None, :] * N[None, :, :]).reshape(C * C, -1)
    return np.asarray(cosine_similarity(sp.csr_matrix(S).T, sp.csr_matrix(Q), dense_output=True))


def cosg_regularise(cos, mu):
    """lambda*cos with lambda = cos^2/((1-mu)cos^2 + mu*sum_groups cos^2); laris/tools/_utils.py:869-880.

    Entries whose cosine is exactly 0 stay 0 (the reference works on sparse .data).
    """
    sq = cos * cos
    tot = sq.sum(1, keepdims=True)
    den = tot if mu == 1 else (1 - mu) * sq + mu * tot
    out = np.zeros_like(cos)
    np.divide(sq, den, out=out, where=cos != 0)
    return out * cos


def cosg_scores(X, labels, C, mu, expressed_pct, remove_lowly_expressed):
    """COSG gene x group score (before indexByGene/iqrLogNormalize).

    cosg.cosg (cosg>=1.0.3, un-vendored; call site laris/tools/_utils.py:580-588)
    applies the same cosine + regularisation that the reference clones in
    laris/tools/_utils.py:861-880, with the low-expression mask of the commented
    block :919-923 (score -> -1).  PARITY UNPINNED: restated from the published
    COSG algorithm, no reference test or source available.
    """
    from sklearn.metrics.pairwise import cosine_similarity
    n = X.shape[0]
    onehot = sp.csr_matrix((np.ones(n), (labels, np.arange(n))), shape=(C, n))
    cos = np.asarray(cosine_similarity(sp.csr_matrix(X).T, onehot, dense_output=True))
    score = cosg_regularise(cos, mu)
    if remove_lowly_expressed:
        Xb = sp.csr_matrix(X, copy=True)
        Xb.eliminate_zeros()
        Xb.data = np.ones_like(Xb.data)
        cnt = np.asarray((onehot @ Xb).todense()).T               # G x C
        n_c = np.bincount(labels, minlength=C)
        score[cnt < n_c[None, :] * expressed_pct] = -1.0
    return score


# ---------------------------------------------------------------------------
# a11  interaction score table
# ---------------------------------------------------------------------------


def interaction_table(spec_l, spec_r, spatial_score, frac, spatial_weight):
    """I[p,s,r] = spec_l[p,s]*spec_r[p,r]*score_p^w * frac[p,s]*frac[p,r].

    laris/tools/_utils.py:1351-1406.  ``spec_l/spec_r`` are the P x C rows of
    the gene specificity table for each pair's ligand / receptor.
    """
    factor = np.ones_like(spatial_score) if spatial_weight == 0 else np.power(spatial_score, spatial_weight)
    inter = spec_l[:, :, None] * spec_r[:, None, :] * factor[:, None, None]
    return (frac[:, :, None] * frac[:, None, :]) * inter


def rescale_top100(flat_scores):
    """x 0.1/mean(top 100) if that mean > 0; laris/tools/_utils.py:1426-1437."""
    top = np.sort(flat_scores)[::-1][: min(100, flat_scores.size)]
    m = top.mean() if top.size else 0.0
    return (0.1 / m) if m > 0 else 1.0


# ---------------------------------------------------------------------------
# a12-a14  background sets, permutation p-values, BH-FDR
# ---------------------------------------------------------------------------


def background_interactions(S, nb, leaf_size=30):
    """30-NN of each pair in (mean, variance) space; laris/tools/_utils.py:1075-1105."""
    from sklearn.neighbors import KDTree
    S = sp.csr_matrix(S)
    mean = np.asarray(S.mean(axis=0)).ravel()
    var = np.asarray(S.power(2).mean(axis=0)).ravel() - mean ** 2
    feats = np.column_stack([mean, var])
    _, idx = KDTree(feats, leaf_size=leaf_size, metric="euclidean").query(feats, k=nb + 1)
    return idx[:, 1:].astype(np.int64), feats


def pvalue_counts(table, in_table, bg, draws_of_row):
    """Exceedance counts of the background-resampling null.

    laris/tools/_utils.py:1518-1596.  ``table[g, p]`` is the score of pair p in
    sender-receiver group g (only meaningful where ``in_table[p]``); the null
    value of draw m for row (g,p) is ``table[g, bg[p, m]]`` or 0.0 when that
    background pair is not in the table (:1563).  ``draws_of_row(g, p_idx)``
    returns int draws [len(p_idx), N].
    Returns (ge[g,p] = #{null >= obs}, nz[g,p] = #{null > 0}, ge_nz = #{null >= obs and null > 0}).
    """
    Gn, P = table.shape
    ge = np.zeros((Gn, P), np.int64)
    nz = np.zeros((Gn, P), np.int64)
    ge_nz = np.zeros((Gn, P), np.int64)
    pidx = np.flatnonzero(in_table)
    masked = np.where(in_table[None, :], table, 0.0)
    for g in range(Gn):
        d = draws_of_row(g, pidx)
        null = masked[g][bg[pidx][np.arange(len(pidx))[:, None], d]]
        obs = table[g, pidx][:, None]
        ge[g, pidx] = (null >= obs).sum(1)
        nz[g, pidx] = (null > 0).sum(1)
        ge_nz[g, pidx] = ((null >= obs) & (null > 0)).sum(1)
    return ge, nz, ge_nz


def pvalues_from_counts(table, ge, nz, ge_nz, n_perm, conditional):
    """laris/tools/_utils.py:1573-1590 (conditional) / :1591-1594 (standard)."""
    if not conditional:
        return (ge + 1) / (n_perm + 1)
    p = np.ones(table.shape)
    pos = table > 0
    p[pos] = (ge_nz[pos] + 1) / (nz[pos] + 1)
    return p


def bh_fdr(p):
    """Benjamini-Hochberg adjusted p (statsmodels multipletests 'fdr_bh',
    call site laris/tools/_utils.py:1636-1646; statsmodels>=0.13.2 un-vendored,
    textbook algorithm)."""
    p = np.asarray(p, dtype=np.float64)
    m = p.size
    if m == 0:
        return p.copy()
    order = np.argsort(p, kind="stable")
    adj = np.minimum.accumulate((p[order] * m / np.arange(1, m + 1))[::-1])[::-1]
    out = np.empty(m)
    out[order] = np.minimum(adj, 1.0)
    return out


def fdr_by_group(pvals, scores, prefilter_fdr, prefilter_threshold):
    """Per-group BH on p.round(6) of rows with score > threshold, others 1.0;
    clip <= 1, NaN -> 1, -log10(fdr + 1e-10) clipped >= 0.
    laris/tools/_utils.py:1617-1658.  Inputs are [groups, rows]."""
    fdr = np.ones_like(pvals)
    pr = np.round(pvals, 6)
    for g in range(pvals.shape[0]):
        sel = scores[g] > prefilter_threshold if prefilter_fdr else np.ones(pvals.shape[1], bool)
        if sel.any():
            fdr[g, sel] = bh_fdr(pr[g, sel])
    fdr = np.nan_to_num(np.minimum(fdr, 1.0), nan=1.0)
    return fdr, np.maximum(-np.log10(fdr + 1e-10), 0.0)


# ---------------------------------------------------------------------------
# extension: label-permutation null of the co-localisation cosines (no reference counterpart; the reference's nulls
# are the random graph, laris/tools/_utils.py:436-446, and the background-pair resampling, :1544-1570)
# ---------------------------------------------------------------------------

LABEL_PERM_BASE = 0x40000000          # key stream of permutation t: LABEL_PERM_BASE + t (csrc/rng.cuh kLabelPermBase)


def label_permutation_null(S, labels, ind, w, C, n_perm, seed):
    """Cell-type labels are permuted over the cells (a keyed bijection of [0, n): oracle.rng.feistel_permutation), the
    neighbourhood composition (:839-856) and the P x C^2 cosines (:859-866) are recomputed, and
    ge[p,(a,b)] = #{t : cos_t >= cos_observed}.  Returns (cos_observed, ge, p = (ge + 1) / (n_perm + 1))."""
    from . import rng as orng
    labels = np.asarray(labels)
    n = len(labels)
    cos0 = celltype_pair_cosine(S, neighborhood_composition(labels, ind, w, C))
    ge = np.zeros(cos0.shape, dtype=np.int64)
    for t in range(n_perm):
        lt = labels[orng.feistel_permutation(n, LABEL_PERM_BASE + t, seed)]
        ge += celltype_pair_cosine(S, neighborhood_composition(lt, ind, w, C)) >= cos0
    return cos0, ge, (ge + 1.0) / (n_perm + 1.0)


# ---------------------------------------------------------------------------
# whole-path drivers (chain the pieces in the reference's order)
# ---------------------------------------------------------------------------


def run_step1(S, xy, k, sigma, n_repeats, random_seed, mu, thr, n_top_lr, mode="numpy"):
    """runLARIS step 1, laris/tools/__init__.py:459-526.  Returns a dict."""
    ind, dist = knn_graph(xy, k, include_self=False)
    w = decay_weights(dist, sigma)
    gsp, rand, score = spatial_specificity(S, ind, w, n_repeats, random_seed, mu, mode)
    Sz = sp.csr_matrix(S, copy=True)
    Sz.eliminate_zeros()
    nnz = np.asarray(Sz.getnnz(axis=0)).ravel()
    order, ranked = rank_pairs(score, nnz, thr, n_top_lr)
    return dict(gsp=gsp, rand=rand, score=score, nnz=nnz, order=order, ranked=ranked)


def run_step2(X, S, xy, labels, C, lig_gene, rec_gene, top_order, top_score, *, k_ct, mu_ct,
              expressed_pct, remove_lowly_expressed, nb, n_perm, spatial_weight, score_threshold,
              prefilter_fdr, prefilter_threshold, conditional, chunk_size, draw_fn, normalise):
    """_calculate_laris_score_by_celltype, laris/tools/_utils.py:1306-1658.

    ``lig_gene/rec_gene``: gene column (into X) of each of the P pairs.
    ``top_order/top_score``: step-1 ranking (pairs in laris_lr order, their scores).
    ``draw_fn(group_index, n_rows)`` -> int draws [n_rows, n_perm] for the rows of
    that sender-receiver group in table order (so the RNG is injected).
    ``normalise(table, row_names?)`` stands for cosg.indexByGene+iqrLogNormalize
    (parity unpinned): applied to the gene x C specificity and the P x C^2
    neighbourhood table.
    Returns dict with the long table in the reference's final order.
    """
    P = S.shape[1]
    # :1306  gene specificity over the unique LR genes of the ranked pairs
    genes = np.unique(np.concatenate([lig_gene[top_order], rec_gene[top_order]]))
    spec = cosg_scores(sp.csc_matrix(X)[:, genes], labels, C, mu_ct, expressed_pct, remove_lowly_expressed)
    spec = normalise(np.where(spec == -1, 0.0, spec))
    pos = {g: i for i, g in enumerate(genes)}
    sl = spec[[pos[g] for g in lig_gene[top_order]]]
    sr = spec[[pos[g] for g in rec_gene[top_order]]]
    # :1325-1337
    frac = celltype_fraction(S, labels, C)[top_order]
    # :1351-1406  melt order: pair-major, then receiver-major, sender fastest
    inter = interaction_table(sl, sr, top_score, frac, spatial_weight)       # [pair, s, r]
    flat = inter.transpose(0, 2, 1).reshape(-1)                              # [pair, r, s]
    npair = len(top_order)
    pair_of = np.repeat(np.arange(npair), C * C)
    recv_of = np.tile(np.repeat(np.arange(C), C), npair)
    send_of = np.tile(np.arange(C), npair * C)
    # :1422-1437
    o1 = np.argsort(-flat, kind="stable")
    flat, pair_of, recv_of, send_of = flat[o1], pair_of[o1], recv_of[o1], send_of[o1]
    flat = flat * rescale_top100(flat)
    # :1448-1472  co-localisation weight
    ind, dist = knn_graph(xy, k_ct, include_self=False)
    N = neighborhood_composition(labels, ind, decay_weights(dist), C)
    ctc_raw = cosg_regularise(celltype_pair_cosine(S, N), mu_ct)            # P x C^2 (sender-major)
    ctc = normalise(ctc_raw)
    flat = flat * ctc[top_order[pair_of], send_of * C + recv_of]
    o2 = np.argsort(-flat, kind="stable")
    flat, pair_of, recv_of, send_of = flat[o2], pair_of[o2], recv_of[o2], send_of[o2]
    # :1480-1489
    flat = flat.astype(np.float64)
    flat[flat < score_threshold] = 0.0
    out = dict(score=flat, pair=top_order[pair_of], sender=send_of, receiver=recv_of,
               ctc_raw=ctc_raw, frac=frac, spec=spec, genes=genes)
    if n_perm <= 0:
        return out
    # :1507  background sets over ALL P pairs
    bg, feats = background_interactions(S, nb)
    out["bg"] = bg
    # :1518-1596  groups in sorted (sender, receiver) order, rows in table order
    pvals = np.full(flat.shape, np.nan)
    group_id = send_of * C + recv_of
    for g in np.unique(group_id):
        rows = np.flatnonzero(group_id == g)
        lookup = np.zeros(P)
        present = np.zeros(P, bool)
        lookup[out["pair"][rows]] = flat[rows]
        present[out["pair"][rows]] = True
        for s in range(0, len(rows), chunk_size):
            rr = rows[s:s + chunk_size]
            d = draw_fn(int(g), len(rr))
            ctrl = bg[out["pair"][rr]][np.arange(len(rr))[:, None], d]
            null = np.where(present[ctrl], lookup[ctrl], 0.0)
            obs = flat[rr][:, None]
            if conditional:
                pv = np.ones(len(rr))
                m = obs[:, 0] > 0
                nzm = null > 0
                pv[m] = ((((null >= obs) & nzm).sum(1) + 1) / (nzm.sum(1) + 1))[m]
            else:
                pv = ((null >= obs).sum(1) + 1) / (n_perm + 1)
            pvals[rr] = pv
    out["p_value"] = pvals
    # :1617-1658
    fdr = np.full(flat.shape, np.nan)
    for g in np.unique(group_id):
        rows = np.flatnonzero(group_id == g)
        pr = np.round(pvals[rows], 6)
        sel = flat[rows] > prefilter_threshold if prefilter_fdr else np.ones(len(rows), bool)
        f = np.ones(len(rows))
        if sel.any():
            f[sel] = bh_fdr(pr[sel])
        fdr[rows] = f
    fdr = np.nan_to_num(np.minimum(fdr, 1.0), nan=1.0)
    out["p_value_fdr"] = fdr
    out["nlog10_p_value_fdr"] = np.maximum(-np.log10(fdr + 1e-10), 0.0)
    return out
