"""CPU tier: the numpy restatement (oracle.restate) against the fixtures written
by the UNMODIFIED reference (oracle/make_golden.py), plus the reference's two
docstring known answers."""
import os

import numpy as np
import pandas as pd
import scipy.sparse as sp

from oracle import cosg_standin
from oracle import restate as R
from oracle import rng as orng

GOLD = os.path.join(os.path.dirname(__file__), "golden")
P_ = dict(k_prepare=8, k_step1=8, k_ct=8, sigma=100.0, R=3, seed=27, mu=1.0, mu_ct=100.0, thr=20, nb=30, N=200,
          expressed_pct=0.1)


def _inputs(golden):
    X = sp.csr_matrix((golden["X_data"], golden["X_indices"], golden["X_indptr"]), shape=tuple(golden["X_shape"]))
    return X, golden["xy"], golden["labels"], golden["lig_idx"], golden["rec_idx"]


def test_fixture_matches_generator(golden, tiny):
    a, lr, c = tiny
    X, xy, labels, li, ri = _inputs(golden)
    assert np.array_equal(a.obsm["X_spatial"], xy)
    assert (sp.csr_matrix(a.X) != X).nnz == 0
    assert np.array_equal(a.obs["CellTypes"].cat.codes.to_numpy(), labels)
    assert np.array_equal(R.gene_index(a.var_names, lr["ligand"]), li)


def test_knn_matches_reference(golden):
    _, xy, *_ = _inputs(golden)
    ind, dist = R.knn_graph(xy, P_["k_prepare"], True)
    assert np.array_equal(ind, golden["knn_self_idx"]) and np.array_equal(dist, golden["knn_self_dist"])
    ind2, dist2 = R.knn_graph(xy, P_["k_step1"], False)
    assert np.array_equal(ind2, golden["knn_noself_idx"])
    assert np.allclose(R.decay_weights(dist2, P_["sigma"]), golden["knn_noself_w"], rtol=1e-15)
    # arithmetic restatement of the KD-tree result (tie rule, exclude-self rule)
    bi, bd = R.knn_bruteforce(xy, P_["k_prepare"], True)
    assert np.array_equal(bi, ind) and np.array_equal(bd, dist)
    bi, bd = R.knn_bruteforce(xy, P_["k_step1"], False)
    assert np.array_equal(bi, ind2) and np.array_equal(bd, dist2)


def test_prepare_matches_reference(golden):
    X, xy, _, li, ri = _inputs(golden)
    ind, dist = R.knn_graph(xy, P_["k_prepare"], True)
    S = R.lr_scores(X, ind, R.decay_weights(dist), li, ri)
    assert np.array_equal(S.indptr, golden["S_indptr"]) and np.array_equal(S.indices, golden["S_indices"])
    assert np.allclose(S.data, golden["S_data"], rtol=1e-14, atol=0)


def _scores(golden):
    n, P = golden["xy"].shape[0], len(golden["lig_idx"])
    return sp.csr_matrix((golden["S_data"], golden["S_indices"], golden["S_indptr"]), shape=(n, P))


def test_step1_matches_reference(golden):
    S = _scores(golden)
    s1 = R.run_step1(S, golden["xy"], P_["k_step1"], P_["sigma"], P_["R"], P_["seed"], P_["mu"], P_["thr"], S.shape[1])
    assert np.allclose(s1["gsp"], golden["gsp"], rtol=1e-12, atol=1e-15)
    assert np.allclose(s1["rand"], golden["rand"], rtol=1e-12, atol=1e-15)
    assert np.array_equal(s1["nnz"], golden["nnz"])
    assert np.array_equal(s1["order"], golden["rank_order"])
    assert np.allclose(s1["ranked"], golden["rank_score"], rtol=1e-12, atol=1e-15)


def test_step2_matches_reference(golden):
    X, xy, labels, li, ri = _inputs(golden)
    S = _scores(golden)
    C = int(labels.max()) + 1
    seeds = orng.numpy_repeat_seeds(P_["seed"], P_["R"])
    rs = orng.numpy_null_graph(xy.shape[0], P_["k_step1"], seeds[-1])[2]
    s2 = R.run_step2(X, S, xy, labels, C, li, ri, golden["rank_order"], golden["rank_score"], k_ct=P_["k_ct"],
                     mu_ct=P_["mu_ct"], expressed_pct=P_["expressed_pct"], remove_lowly_expressed=True, nb=P_["nb"],
                     n_perm=P_["N"], spatial_weight=1.0, score_threshold=1e-6, prefilter_fdr=True,
                     prefilter_threshold=0.0, conditional=False, chunk_size=50000,
                     draw_fn=lambda g, rows: rs.randint(0, P_["nb"], size=(rows, P_["N"])),
                     normalise=lambda M: cosg_standin.iqr_log_normalize(np.asarray(M)))
    assert np.allclose(s2["ctc_raw"], golden["ctc_raw"], rtol=1e-12, atol=1e-18)
    assert np.array_equal(s2["bg"], golden["bg"])
    key = (s2["pair"] * C + s2["sender"]) * C + s2["receiver"]
    o = np.argsort(key)
    assert np.array_equal(key[o], golden["tbl_key"])
    assert np.allclose(s2["score"][o], golden["tbl_score"], rtol=1e-9, atol=1e-15)
    assert np.allclose(s2["p_value"][o], golden["tbl_p"], rtol=1e-12)            # permutation counts: exact
    assert np.allclose(s2["p_value_fdr"][o], golden["tbl_fdr"], rtol=1e-9)
    assert np.allclose(s2["nlog10_p_value_fdr"][o], golden["tbl_nlog"], rtol=1e-9, atol=1e-12)
    # row ORDER of the reference's table (melt, sort by the pre-weight score, sort by the final score - unstable sorts, so
    # only rows with a distinct positive score are comparable): the restatement's stable sorts put them in the same rows
    pos = golden["tbl_score"] > 0
    distinct = pos & (np.unique(golden["tbl_score"], return_counts=True)[1][np.searchsorted(np.unique(golden["tbl_score"]),
                                                                                             golden["tbl_score"])] == 1)
    assert distinct.sum() > 400
    assert golden["tbl_pos"][pos].max() == pos.sum() - 1 and o[pos].max() == pos.sum() - 1      # positive rows come first
    assert np.array_equal(golden["tbl_pos"][distinct], o[distinct])


def test_docstring_known_answers():
    kat = np.load(os.path.join(GOLD, "docstring_kat.npz"))
    # laris/tools/_utils.py:133-135
    assert list(R.select_top_n(np.array([0.1, 0.9, 0.3, 0.7, 0.5]), 3)) == [1, 3, 4] == list(kat["top_n"])
    # laris/tools/_utils.py:188-197
    m = np.array([[1., 0, 2, 0, 1], [0, 1, 1, 0, 2], [2, 0, 0, 1, 1]])
    prod = (m[:, None, :] * m[None, :, :]).reshape(9, 5)
    assert np.array_equal(prod, kat["pair_rows"])
    assert list(kat["pair_names"][:3]) == ["TypeA::TypeA", "TypeA::TypeB", "TypeA::TypeC"]


def test_bh_fdr_textbook():
    p = np.array([0.01, 0.04, 0.03, 0.005, 0.5, 0.5])
    adj = R.bh_fdr(p)
    order = np.argsort(p)
    expect = np.minimum.accumulate((p[order] * 6 / np.arange(1, 7))[::-1])[::-1]
    assert np.allclose(adj[order], np.minimum(expect, 1))
    assert R.bh_fdr(np.array([])).size == 0


def test_philox_mode_null_is_a_valid_null(golden):
    """Same statistic under the device RNG twin: finite, in [-1,1], close in
    distribution to the numpy-stream null (both are uniform random graphs)."""
    S = _scores(golden)
    ind, dist = R.knn_graph(golden["xy"], P_["k_step1"], False)
    w = R.decay_weights(dist, P_["sigma"])
    _, rand_p, _ = R.spatial_specificity(S, ind, w, 3, 27, 1.0, "philox")
    assert np.isfinite(rand_p).all() and (np.abs(rand_p) <= 1 + 1e-12).all()
    assert abs(rand_p.mean() - golden["rand"].mean()) < 0.02


def test_radius_graph_oracle_is_the_plain_definition():
    """Extension oracle (K1r): sklearn's radius graph equals the brute-force definition d <= r on tie-free points."""
    from oracle import restate as R
    rng = np.random.default_rng(5)
    xy = rng.random((400, 2)) * 50.0
    for include_self in (True, False):
        indptr, indices, dist = R.radius_graph(xy, 6.0, include_self)
        d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
        adj = d <= 6.0
        if not include_self:
            np.fill_diagonal(adj, False)
        assert np.array_equal(np.diff(indptr), adj.sum(1))
        rows = np.repeat(np.arange(400), np.diff(indptr))
        assert adj[rows, indices].all() and np.allclose(dist, d[rows, indices], rtol=1e-12, atol=1e-12)
