"""Generate tests/golden/*.npz from the UNMODIFIED reference.  Build container only.

    python -m oracle.make_golden

Runs the reference's own ``laris.tools`` (loaded under stub modules by
oracle.ref_loader) on a small seeded dataset, checks the numpy
restatement (oracle.restate) against it, and writes the reference's outputs as
fixtures.  The GPU box has no /root/reference, so tests there compare against
these fixtures and against oracle.restate (which this script pins).

``runLARIS`` itself cannot run under pandas 3 (read-only ``.values`` at
laris/tools/__init__.py:505-510); its step-1 glue lines are
replayed here by calling the same ``_utils`` helpers in the same order so the
global numpy RNG stream is consumed exactly as the reference consumes it.
"""
from __future__ import annotations

import contextlib
import io
import os
import warnings

import numpy as np
import pandas as pd
import scipy.sparse as sp

from laris_b200.synth import make_dataset
from oracle import cosg_standin
from oracle import restate as R
from oracle import rng as orng
from oracle.ref_loader import load_reference_tools

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")

TINY = dict(cfg=1, n=1500, G=300, G_u=160, P=120, C=5)
PARAMS = dict(k_prepare=8, k_step1=8, k_ct=8, sigma=100.0, R=3, seed=27, mu=1.0, mu_ct=100.0,
              thr=20, nb=30, N=200, expressed_pct=0.1)


def normalise(M):
    return cosg_standin.iqr_log_normalize(np.asarray(M))


def main():
    T = load_reference_tools()
    U = T._utils
    a, lr, c = make_dataset(**TINY)
    p = PARAMS
    xy = a.obsm["X_spatial"]
    n = a.n_obs
    quiet = io.StringIO()

    # ---- reference: prepareLRInteraction (unmodified) -----------------
    ref = T.prepareLRInteraction(a, lr, number_nearest_neighbors=p["k_prepare"])
    Sref = sp.csr_matrix(ref.X)
    Sref.sort_indices()

    # ---- reference: runLARIS step 1 via its own helpers ----------------
    A = U._build_adjacency_matrix(ref, use_rep="X_spatial", n_nearest_neighbors=p["k_step1"], sigma=p["sigma"])
    gx = ref.X.T
    gsp_ref = np.asarray(U._rowwise_cosine_similarity(gx, gx @ A.T)).ravel()
    rand_list = U._generate_random_background(ref, A, gx, n_nearest_neighbors=p["k_step1"],
                                              n_repeats=p["R"], random_seed=p["seed"])
    rand_ref = np.mean(rand_list, axis=0)
    rand_ref = np.asarray(rand_ref).ravel()
    score_ref = gsp_ref - p["mu"] * rand_ref
    # glue laris/tools/__init__.py:493-526 with writable copies
    nnz_ref = np.asarray(Sref.getnnz(axis=0)).ravel()
    var = pd.DataFrame({"s": score_ref, "nnz": nnz_ref}, index=ref.var_names).sort_values("s", ascending=False)
    rk = var["s"].to_numpy().copy()
    rk[var["nnz"].to_numpy() < p["thr"]] = rk.min() - 0.001
    top = U._select_top_n(rk, len(rk))
    names = var.index.to_numpy()[top]
    laris_lr = pd.DataFrame({"ligand": [x.split("::")[0] for x in names],
                             "receptor": [x.split("::")[1] for x in names],
                             "score": rk[top]}, index=names)
    laris_lr["Rank"] = np.arange(len(names))

    # ---- reference: step 2 (unmodified function; cosg stand-ins, see ref_loader) ----
    with contextlib.redirect_stdout(quiet), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = U._calculate_laris_score_by_celltype(
            adata=a, lr_adata=ref, laris_lr=laris_lr, groupby="CellTypes", use_rep_spatial="X_spatial",
            number_nearest_neighbors=p["k_ct"], mu=p["mu_ct"], expressed_pct=p["expressed_pct"],
            remove_lowly_expressed=True, mask_threshold=1e-6, calculate_pvalues=True, layer=None,
            n_nearest_neighbors=p["nb"], n_permutations=p["N"], chunk_size=50000, prefilter_fdr=True,
            prefilter_threshold=0.0, score_threshold=1e-6, spatial_weight=1.0, use_conditional_pvalue=False)
    res["p_value_fdr"] = res["p_value_fdr"].fillna(1.0)       # intended effect of the no-op at _utils.py:1654
    ctc_raw_ref = cosg_standin.index_by_gene(ref.uns["cosg_lr"]["COSG"]).loc[ref.var_names].to_numpy()
    bg_ref = U._prepare_background_interactions(ref, n_nearest_neighbors=p["nb"]).to_numpy()

    # ---- restatement on the same inputs --------------------------------
    li = R.gene_index(a.var_names, lr["ligand"])
    ri = R.gene_index(a.var_names, lr["receptor"])
    ind1, d1 = R.knn_graph(xy, p["k_prepare"], True)
    S = R.lr_scores(a.X, ind1, R.decay_weights(d1), li, ri)
    assert (S.indptr == Sref.indptr).all() and (S.indices == Sref.indices).all()
    assert np.allclose(S.data, Sref.data, rtol=1e-14, atol=0)
    s1 = R.run_step1(S, xy, p["k_step1"], p["sigma"], p["R"], p["seed"], p["mu"], p["thr"], S.shape[1], "numpy")
    assert np.allclose(s1["gsp"], gsp_ref, rtol=1e-12, atol=1e-15)
    assert np.allclose(s1["rand"], rand_ref, rtol=1e-12, atol=1e-15)
    lr_names = ref.var_names.to_numpy()
    assert (lr_names[s1["order"]] == names).all(), "ranking order differs"
    labels = a.obs["CellTypes"].cat.codes.to_numpy()
    seeds = orng.numpy_repeat_seeds(p["seed"], p["R"])
    rs = orng.numpy_null_graph(n, p["k_step1"], seeds[-1])[2]
    s2 = R.run_step2(a.X, S, xy, labels, c.C, li, ri, s1["order"], s1["ranked"], k_ct=p["k_ct"], mu_ct=p["mu_ct"],
                     expressed_pct=p["expressed_pct"], remove_lowly_expressed=True, nb=p["nb"], n_perm=p["N"],
                     spatial_weight=1.0, score_threshold=1e-6, prefilter_fdr=True, prefilter_threshold=0.0,
                     conditional=False, chunk_size=50000,
                     draw_fn=lambda g, rows: rs.randint(0, p["nb"], size=(rows, p["N"])), normalise=normalise)
    assert np.allclose(s2["ctc_raw"], ctc_raw_ref, rtol=1e-12, atol=1e-18)
    assert (s2["bg"] == bg_ref).all()
    # keyed comparison of the long table (tie order of equal scores is unstable in the reference)
    cats = list(a.obs["CellTypes"].cat.categories)
    pair_id = pd.Index(lr_names).get_indexer(res["interaction_name"].to_numpy())
    key_ref = (pair_id * c.C + pd.Index(cats).get_indexer(res["sender"])) * c.C + pd.Index(cats).get_indexer(res["receiver"])
    key_me = (s2["pair"] * c.C + s2["sender"]) * c.C + s2["receiver"]
    o_ref, o_me = np.argsort(key_ref), np.argsort(key_me)
    assert (key_ref[o_ref] == key_me[o_me]).all()
    for col, mine in (("interaction_score", s2["score"]), ("p_value", s2["p_value"]),
                      ("p_value_fdr", s2["p_value_fdr"]), ("nlog10_p_value_fdr", s2["nlog10_p_value_fdr"])):
        r = res[col].to_numpy()[o_ref]
        m = mine[o_me]
        assert np.allclose(r, m, rtol=1e-9, atol=1e-15), (col, np.abs(r - m).max())
    print("restatement == reference on the tiny case: prepare, step 1, step 2 (table, p-values, FDR)")

    # ---- reference: _compute_avg_expression (laris/tools/_utils.py:248-340, unmodified; all groups / a gene subset -
    # the `groups=` form depends on real AnnData views dropping unused categories, which the duck-typed carrier does not) ----
    avg_gene_names = ["g00003", "g00150", "g00007"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        avg_all_ref = U._compute_avg_expression(a, groupby="CellTypes")
        avg_genes_ref = U._compute_avg_expression(a, groupby="CellTypes", genes=avg_gene_names)
    assert list(avg_all_ref.index) == list(a.obs["CellTypes"].cat.categories) and list(avg_all_ref.columns) == list(a.var_names)

    # ---- fixtures -------------------------------------------------------
    os.makedirs(OUT, exist_ok=True)
    X = sp.csr_matrix(a.X)
    A2 = sp.csr_matrix(A)
    np.savez_compressed(
        os.path.join(OUT, "tiny_reference.npz"),
        params=np.array([repr({**TINY, **PARAMS})]),
        xy=xy, labels=labels, X_data=X.data, X_indices=X.indices, X_indptr=X.indptr, X_shape=np.array(X.shape),
        lig_idx=li, rec_idx=ri,
        S_data=Sref.data, S_indices=Sref.indices, S_indptr=Sref.indptr,
        knn_noself_idx=A2.indices.reshape(n, -1), knn_noself_w=A2.data.reshape(n, -1),
        knn_self_idx=ind1, knn_self_dist=d1,
        gsp=gsp_ref, rand=rand_ref, rand_each=np.asarray([np.asarray(x).ravel() for x in rand_list]),
        rank_order=s1["order"], rank_score=rk[top], nnz=nnz_ref,
        ctc_raw=ctc_raw_ref, bg=bg_ref,
        tbl_key=key_ref[o_ref], tbl_score=res["interaction_score"].to_numpy()[o_ref],
        tbl_p=res["p_value"].to_numpy()[o_ref], tbl_fdr=res["p_value_fdr"].to_numpy()[o_ref],
        tbl_nlog=res["nlog10_p_value_fdr"].to_numpy()[o_ref],
        tbl_pos=o_ref,                     # row of the reference's table that holds the j-th key: its row ORDER
        avg_all=avg_all_ref.to_numpy(dtype=np.float64), avg_genes=avg_genes_ref.to_numpy(dtype=np.float64),
        avg_gene_query=np.array(avg_gene_names, dtype=str), avg_gene_columns=np.array(list(avg_genes_ref.columns), dtype=str),
    )
    # docstring known answers (laris/tools/_utils.py:133-135, :188-197)
    m = sp.csr_matrix(np.array([[1., 0, 2, 0, 1], [0, 1, 1, 0, 2], [2, 0, 0, 1, 1]]))
    prod, nm = U._pairwise_row_multiply(m, ["TypeA", "TypeB", "TypeC"])
    np.savez_compressed(os.path.join(OUT, "docstring_kat.npz"),
                        top_n=U._select_top_n(np.array([0.1, 0.9, 0.3, 0.7, 0.5]), 3),
                        pair_rows=prod.toarray(), pair_names=np.array(nm, dtype=str))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
