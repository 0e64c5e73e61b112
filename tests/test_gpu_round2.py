"""GPU tier, round-2 additions: the device tables of runLARIS step 2 (T1..T5), the batched K4 (TMA-staged observed
tiles, multi-repeat null pass), the key-gather K5, device-resident outputs, cache validation - each against the CPU
oracle / numpy on the same seeded inputs."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import laris_b200._lib  # noqa: F401
    return torch.device("cuda:0")


def _d(x, dtype):
    from laris_b200 import _kernels as K
    return K.to_device(np.ascontiguousarray(x), dtype)


# ------------------------------------------------------------- tables ----
@pytest.mark.parametrize("P,C,mu", [(37, 3, 100.0), (500, 20, 1.0), (2000, 40, 100.0), (5, 1, 0.5)])
def test_pair_cosine_regularise(dev, P, C, mu):
    from laris_b200 import _kernels as K
    from oracle import restate as R
    rng = np.random.default_rng(P + C)
    L = C * C
    num = rng.random((P, L)) * (rng.random((P, L)) < 0.6)
    ss = rng.random(P) + 0.1
    qn2 = rng.random(L) + 0.05
    ss[0] = 0.0                                            # an all-zero pair
    qn2[L - 1] = 0.0                                       # an empty profile
    cos, reg = K.pair_cosine_regularise(_d(num, torch.float64), _d(ss, torch.float64), _d(qn2, torch.float64), mu, want_cos=True)
    den = np.sqrt(ss)[:, None] * np.sqrt(qn2)[None, :]
    ref_cos = np.zeros_like(num)
    np.divide(num, den, out=ref_cos, where=den != 0)
    assert np.allclose(cos.cpu().numpy(), ref_cos, rtol=1e-14, atol=0)
    assert np.allclose(reg.cpu().numpy(), R.cosg_regularise(ref_cos, mu), rtol=1e-12, atol=0)


@pytest.mark.parametrize("rows,cols", [(2000, 1600), (120, 25), (7, 3), (9001, 5)])
def test_column_iqr_log_matches_the_oracle_stand_in(dev, rows, cols):
    from laris_b200 import _kernels as K
    from oracle import cosg_standin
    rng = np.random.default_rng(rows)
    t = rng.random((rows, cols)) ** 4
    t[rng.random((rows, cols)) < 0.5] = 0.0
    t[rng.random((rows, cols)) < 0.05] *= -1.0
    t[:, 0] = 0.0                                          # no positive entry
    if cols > 2:
        t[:, 1] = np.where(np.arange(rows) % 3 == 0, 0.25, 0.0)       # all positives equal: IQR = 0 -> scale = max
        t[:, 2] = 0.0
        t[rows // 2, 2] = 3.0                              # a single positive entry
    got = K.column_iqr_log(_d(t, torch.float64)).cpu().numpy()
    ref = cosg_standin.iqr_log_normalize(t)
    assert np.allclose(got, ref, rtol=1e-12, atol=0)
    assert got.min() == 0.0 and got.max() == 1.0


@pytest.mark.parametrize("m,k", [(3_200_000, 100), (50, 100), (100, 100), (16385, 7), (1, 1), (40000, 4096)])
def test_topk_mean(dev, m, k):
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(m)
    v = rng.random(m) ** 8
    v[rng.random(m) < 0.3] = 0.0
    if m > 1000:
        v[:300] = v.max()                                  # ties at the top, crossing the cut
    out = K.topk_mean(_d(v, torch.float64), k).cpu().numpy()
    top = np.sort(v)[::-1][: min(k, m)]
    assert np.isclose(out[0], top.mean(), rtol=1e-13, atol=0)
    assert np.isclose(out[1], 0.1 / top.mean() if top.mean() > 0 else 1.0, rtol=1e-13)
    z = K.topk_mean(_d(np.zeros(500), torch.float64), 100).cpu().numpy()
    assert z[0] == 0.0 and z[1] == 1.0


def test_interaction_tables_match_the_oracle(dev):
    """T3a + T3b + T3c + T5 + K7 chained on random inputs == the oracle's table arithmetic."""
    from laris_b200 import _kernels as K
    from oracle import restate as R
    rng = np.random.default_rng(5)
    P, Pt, C, N = 300, 257, 6, 200
    L = C * C
    sl, sr = rng.random((Pt, C)), rng.random((Pt, C))
    sl[rng.random((Pt, C)) < 0.3] = 0.0
    score = rng.random(Pt) + 0.01
    cnt = rng.integers(0, 50, (C, P))
    n_c = rng.integers(50, 90, C).astype(np.float64)
    pair_id = rng.permutation(P)[:Pt]
    ctc = rng.random((P, L)) * (rng.random((P, L)) < 0.7)
    pid_d = _d(pair_id.astype(np.int32), torch.int32)
    inter = K.interaction_scores(_d(sl, torch.float64), _d(sr, torch.float64), _d(score, torch.float64),
                                 _d(cnt, torch.int64), _d(n_c, torch.float64), pid_d, P)
    frac = (cnt / n_c[:, None]).T
    ref_inter = R.interaction_table(sl, sr, score, frac[pair_id], 1.0).reshape(Pt, L)
    assert np.array_equal(inter.cpu().numpy(), ref_inter)                     # same operation order: same bits
    resc = K.topk_mean(inter.view(-1), 100)
    ref_scale = R.rescale_top100(ref_inter.reshape(-1))
    assert np.isclose(resc.cpu().numpy()[1], ref_scale, rtol=1e-13)
    flat, table, active = K.finalize_scores(inter, resc, _d(ctc, torch.float64), pid_d, P, 1e-3)
    ref_flat = ref_inter * resc.cpu().numpy()[1] * ctc[pair_id]
    ref_flat[ref_flat < 1e-3] = 0.0
    assert np.array_equal(flat.cpu().numpy(), ref_flat)
    ref_table = np.zeros((L, P))
    ref_table[:, pair_id] = ref_flat.T
    assert np.array_equal(table.cpu().numpy(), ref_table)
    ref_active = np.zeros(P, bool)
    ref_active[pair_id] = True
    assert np.array_equal(active.cpu().numpy().astype(bool), ref_active)
    # counts -> p -> BH, both p-value definitions
    bg = np.stack([rng.permutation(P)[:30] for _ in range(P)]).astype(np.int32)
    for conditional in (False, True):
        ge, nz, both = K.pvalue_counts(table, _d(bg, torch.int32), N, active=active, seed=3, only_ge=not conditional)
        p, pr, sel = K.pvalues_from_counts(table, active, ge, nz, both, N, conditional, True, 0.0)
        ref_p = R.pvalues_from_counts(ref_table, ge.cpu().numpy(), None if nz is None else nz.cpu().numpy(),
                                      None if both is None else both.cpu().numpy(), N, conditional)
        assert np.array_equal(p.cpu().numpy(), ref_p)
        assert np.array_equal(pr.cpu().numpy(), np.round(ref_p, 6))
        assert np.array_equal(sel.cpu().numpy().astype(bool), ref_active[None, :] & (ref_table > 0.0))
        fdr = K.bh_fdr(pr, sel)
        p_rows, f_rows, l_rows = (x.cpu().numpy() for x in K.gather_result_rows(p, fdr, pid_d))
        ref_fdr, ref_nlog = R.fdr_by_group(ref_p[:, pair_id], ref_table[:, pair_id], True, 0.0)
        assert np.array_equal(p_rows, ref_p[:, pair_id].T)
        assert np.allclose(f_rows, ref_fdr.T, rtol=1e-12, atol=0)
        assert np.allclose(l_rows, ref_nlog.T, rtol=1e-9, atol=1e-12)


# ----------------------------------------------------------------- K4 ----
def _uniform(n, seed):
    return np.random.default_rng(seed).random((n, 2)) * 10.0 * np.sqrt(n)


@pytest.mark.parametrize("n,P,k,use_order", [(3001, 131, 13, True), (20000, 500, 10, True), (1500, 64, 8, False),
                                              (129, 7, 3, True), (5000, 2000, 15, True), (4000, 40, 20, True)])
def test_batched_stats_equal_legacy_and_oracle(dev, n, P, k, use_order):
    """Tiled observed pass (k <= 15; legacy path above) and the multi-repeat null pass against the single-graph kernel
    and the oracle; n % 128 != 0, column tails, and a processing order that is not spatial (tiles overflow their
    staging capacity and read global rows directly)."""
    from laris_b200 import _kernels as K, engine
    from oracle import restate as R
    xy = _uniform(n, 3)
    S = sp.random(n, P, 0.25, random_state=7, format="csr", dtype=np.float32)
    sm = engine.scores_from_csr(S)
    g = engine.build_graph(xy, k, False, sigma=40.0)
    order = g.order if use_order else None
    legacy = K.spatial_stats(sm.S, P, g.idx, g.w, False, g.order).cpu().numpy()
    R_ = 4
    cols, wts, _ = engine.null_graphs(g, R_, 11, "philox")
    both = K.spatial_stats_batched(sm.S, P, g.idx, g.w, order, cols, wts).cpu().numpy()
    assert both.shape == (1 + R_, 5, P)
    assert np.allclose(both[0], legacy, rtol=1e-6, atol=1e-9)                # fp32 tile partials vs fp32 products
    assert np.array_equal(both[0][4], legacy[4]) and np.allclose(both[0][3], legacy[3], rtol=1e-6)
    ind, dist = R.knn_graph(xy, k, False)
    ref = R.observed_cosine(S.astype(np.float64), ind, R.decay_weights(dist, 40.0))
    assert np.allclose(engine._cosine(both[0]), ref, rtol=RTOL, atol=1e-9)
    for r in range(R_):
        one = K.spatial_stats(sm.S, P, cols[r].contiguous(), wts[r].contiguous(), True, None).cpu().numpy()
        # the column-blocked (L2-resident) null pass keeps fp32 partial sums over <= 64 rows per thread
        assert np.allclose(both[1 + r], one, rtol=2e-6, atol=1e-300)
        assert np.array_equal(both[1 + r][4], one[4])
    only_null = K.spatial_stats_batched(sm.S, P, None, None, None, cols, wts).cpu().numpy()
    assert np.array_equal(only_null, both[1:])
    # a scrambled processing order: tiles are no longer local, the union of neighbour rows exceeds the staging cap
    perm = torch.randperm(n, device=dev).to(torch.int32)
    scr = K.spatial_stats_batched(sm.S, P, g.idx, g.w, perm).cpu().numpy()[0]
    assert np.allclose(scr, both[0], rtol=1e-6, atol=1e-9) and np.array_equal(scr[4], legacy[4])


@pytest.mark.parametrize("mode", ["bulk", "gather4", "cpasync"])
def test_tile_staging_modes_agree(dev, mode):
    """The three ways the tiled K4a brings neighbour rows into shared memory (1-D bulk copies, TMA tile::gather4 through a
    tensor map, cp.async) give the legacy kernel's numbers; the mode is fixed per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LARIS_B200_K4A_STAGING=mode)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "time_k4.py"), "30001", "131", "10"], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert f"staging={mode}" in out.stdout and "equal=True" in out.stdout, out.stdout


def test_observed_stats_on_padded_radius_rows(dev):
    """Fixed-degree rows padded with (own row, weight 0) edges - the radius-graph layout - through the tiled pass."""
    from laris_b200 import _kernels as K, engine
    n, P = 6000, 90
    xy = _uniform(n, 5)
    S = sp.random(n, P, 0.3, random_state=2, format="csr", dtype=np.float32)
    sm = engine.scores_from_csr(S)
    g = engine.build_radius_graph(xy, 10.0, False, sigma=30.0)
    assert 4 <= g.k <= 15, g.k
    a = K.spatial_stats(sm.S, P, g.idx, g.w, False, g.order).cpu().numpy()
    b = K.spatial_stats_batched(sm.S, P, g.idx, g.w, g.order).cpu().numpy()[0]
    assert np.allclose(a, b, rtol=1e-6, atol=1e-9)


# ----------------------------------------------------------------- K5 ----
@pytest.mark.parametrize("n,P,C,k,coherent", [(5000, 131, 7, 10, True), (20000, 500, 40, 10, True), (3000, 33, 1, 6, False),
                                              (4097, 64, 64, 12, False), (257, 5, 3, 4, False), (30000, 2000, 20, 10, True)])
def test_key_gather_contraction(dev, n, P, C, k, coherent):
    from laris_b200 import _kernels as K, engine
    rng = np.random.default_rng(n + C)
    xy = _uniform(n, 1)
    if coherent:
        labels = (np.floor(xy[:, 0] / xy[:, 0].max() * 0.999 * C)).astype(np.int32)
    else:
        labels = rng.integers(0, C, n).astype(np.int32)
    S = sp.random(n, P, 0.2, random_state=4, format="csr", dtype=np.float32)
    sm = engine.scores_from_csr(S)
    g = engine.build_graph(xy, k, False, None)
    N = K.neighborhood_composition(_d(labels, torch.int32), g.idx, g.w, C)
    num2, qn2 = K.celltype_pair_contract(sm.S, P, N, 2)
    num1, qn1 = K.celltype_pair_contract(sm.S, P, N, 1)
    Nh = N.cpu().numpy().astype(np.float64)
    Q = (Nh[:, :, None] * Nh[:, None, :]).reshape(n, C * C)
    ref = np.asarray(S.astype(np.float64).T @ Q)
    got = num2.cpu().numpy()
    assert np.allclose(got, ref, rtol=1e-6, atol=1e-9 * np.abs(ref).max())
    assert np.allclose(num1.cpu().numpy(), ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())
    assert np.allclose(qn2.cpu().numpy(), qn1.cpu().numpy(), rtol=1e-12)
    sym = got.reshape(P, C, C)
    assert np.array_equal(sym, sym.transpose(0, 2, 1))
    auto, _ = K.celltype_pair_contract(sm.S, P, N, None, order=g.order)
    assert np.allclose(auto.cpu().numpy(), ref, rtol=2e-5, atol=1e-6 * np.abs(ref).max())


# ------------------------------------------------- device-resident path ----
def test_resident_scores_and_array_output_equal_the_frames(dev):
    """prepareLRInteraction(to_host=False, csr_dev=...) + runLARIS(output='arrays') - the path bench.py times - gives
    the numbers of the default host-to-host calls."""
    import laris_b200 as la
    from laris_b200 import _kernels as K, engine
    from laris_b200._anndata import AnnDataLite
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(1, n=3000, G=200, G_u=100, P=61, C=4)
    kw = dict(n_nearest_neighbors=8, n_repeats=2, number_nearest_neighbors=8, n_permutations=60, n_top_lr=61,
              n_cells_expressed_threshold=10, rng="philox", verbose=False)
    ref_ad = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8)
    ref_lr, ref_tab = la.tl.runLARIS(ref_ad, a, **kw)
    csr_d = engine.upload_csr(a.X)
    a_dev = AnnDataLite(a.X, a.obs, a.var, {"X_spatial": K.to_device(a.obsm["X_spatial"], torch.float64)})
    ad = la.tl.prepareLRInteraction(a_dev, lr, number_nearest_neighbors=8, csr_dev=csr_d, to_host=False)
    assert not sp.issparse(ad.X) and ad.shape == ref_ad.shape
    Xd = ad.X.tocsr()
    assert (Xd != ref_ad.X).nnz == 0
    lr2, arr = la.tl.runLARIS(ad, a_dev, output="arrays", **kw)
    assert list(lr2.index) == list(ref_lr.index) and np.allclose(lr2["score"], ref_lr["score"], rtol=1e-12)
    assert "n_genes_by_counts" not in ad.obs and "cosg_lr" not in ad.uns and "n_genes_by_counts" in ref_ad.obs
    C, names = c.C, np.asarray(ref_ad.var_names)
    pid = pd.Index(names).get_indexer(ref_tab["interaction_name"].to_numpy())
    cats = list(a.obs["CellTypes"].cat.categories)
    g = pd.Index(cats).get_indexer(ref_tab["sender"]) * C + pd.Index(cats).get_indexer(ref_tab["receiver"])
    row_of = np.empty(len(names), np.int64)
    row_of[arr["pair_id"]] = np.arange(len(arr["pair_id"]))
    for col in ("interaction_score", "p_value", "p_value_fdr", "nlog10_p_value_fdr"):
        got = arr[col].cpu().numpy()[row_of[pid], g]
        assert np.allclose(got, ref_tab[col].to_numpy(), rtol=1e-12, atol=0), col


def test_cached_scores_are_revalidated_after_in_place_edits(dev):
    """an in-place edit of lr_adata.X / adata.X between the two calls must not be served from the device cache."""
    import laris_b200 as la
    from laris_b200.synth import make_dataset
    from laris_b200.tools import _utils as U
    a, lr, c = make_dataset(1, n=2500, G=150, G_u=80, P=40, C=3)
    kw = dict(n_nearest_neighbors=8, by_celltype=False, n_cells_expressed_threshold=10, verbose=False)
    ad = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8)
    sm0 = U._scores_of(ad)
    assert U._scores_of(ad) is sm0                                  # untouched: served from the cache
    base = la.tl.runLARIS(ad, None, **kw)
    ad.X.data[:] = np.log1p(ad.X.data)                              # scanpy-style in-place transform
    sm1 = U._scores_of(ad)
    assert sm1 is not sm0
    assert np.allclose(sm1.S[:, :40].cpu().numpy(), ad.X.toarray(), rtol=1e-6)
    edited = la.tl.runLARIS(ad, None, **kw)
    assert not np.allclose(edited["score"].to_numpy(), base["score"].to_numpy())
    fresh = la.tl.runLARIS(la._anndata.AnnDataLite(ad.X.copy(), ad.obs.copy(), pd.DataFrame(index=ad.var_names),
                                                   {"X_spatial": a.obsm["X_spatial"]}), None, **kw)
    assert np.allclose(edited["score"].to_numpy(), fresh["score"].to_numpy(), rtol=1e-12)
    # expression cache: same rule
    genes = U._gene_columns(a, np.unique(np.hstack([lr["ligand"], lr["receptor"]])))
    assert U._expression_of(a.X, genes) is not None
    a.X.data *= 2.0
    assert U._expression_of(a.X, genes) is None


def test_numpy_stream_draws_are_chunked(dev, golden, tiny, monkeypatch):
    """the reference-stream p-value draws are generated a batch of groups at a time; the result does not
    depend on the batch size (and equals the reference fixture)."""
    import laris_b200 as la
    from laris_b200.tools import _utils as U
    a, lr, c = tiny
    kw = dict(n_nearest_neighbors=8, random_seed=27, n_repeats=3, n_cells_expressed_threshold=20, number_nearest_neighbors=8,
              n_permutations=200, rng="numpy", verbose=False)
    res = []
    for nbytes in (256 << 20, 3 * 120 * 200):                       # one batch / three groups per batch
        monkeypatch.setattr(U, "_NUMPY_DRAW_BYTES", nbytes)
        ad = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8)
        _, t = la.tl.runLARIS(ad, a, **kw)
        res.append(t.sort_values(["sender", "receiver", "interaction_name"]).reset_index(drop=True))
    assert np.array_equal(res[0]["p_value"].to_numpy(), res[1]["p_value"].to_numpy())
    assert np.allclose(res[0]["interaction_score"].to_numpy(), res[1]["interaction_score"].to_numpy(), rtol=1e-12)


def test_background_sets_with_duplicate_feature_points(dev):
    """identical (mean, variance) points (e.g. all-zero LR pairs) follow the host KD-tree like the reference."""
    import laris_b200 as la
    from laris_b200._anndata import AnnDataLite
    from laris_b200.tools import _utils as U
    from oracle import restate as R
    rng = np.random.default_rng(0)
    n, P = 800, 90
    S = sp.random(n, P, 0.2, random_state=1, format="csr", dtype=np.float32).tolil()
    S[:, 10:25] = 0.0                                               # fifteen all-zero pairs: identical feature points
    S[:, 30] = S[:, 31]
    S = sp.csr_matrix(S)
    ad = AnnDataLite(S, obs=pd.DataFrame(index=[f"c{i}" for i in range(n)]), var=pd.DataFrame(index=[f"a::b{j}" for j in range(P)]),
                     obsm={"X_spatial": rng.random((n, 2))})
    got = U._prepare_background_interactions(ad, 30).to_numpy()
    ref, _ = R.background_interactions(S.astype(np.float64), 30, leaf_size=30)
    assert np.array_equal(got, ref)


def test_missing_labels_are_rejected(dev, tiny):
    import laris_b200 as la
    a, lr, _ = tiny
    b = a.copy()
    lab = b.obs["CellTypes"].astype(object)
    lab.iloc[3] = None
    b.obs["CellTypes"] = pd.Categorical(lab)
    ad = la.tl.prepareLRInteraction(a, lr.iloc[:10], number_nearest_neighbors=8)
    with pytest.raises(ValueError, match="no cell-type label"):
        la.tl.runLARIS(ad, b, n_nearest_neighbors=8, n_cells_expressed_threshold=1, n_permutations=10, verbose=False)


# ------------------------------------------- label-permutation null ----
def test_permuted_labels_are_the_oracle_bijection(dev):
    from laris_b200 import _kernels as K
    from oracle import restate as R
    from oracle import rng as orng
    n = 40_013
    labels = np.random.default_rng(1).integers(0, 7, n).astype(np.int32)
    lab_d = _d(labels, torch.int32)
    for t in (0, 3, 99):
        got = K.permute_labels(lab_d, t, 27).cpu().numpy()
        perm = orng.feistel_permutation(n, R.LABEL_PERM_BASE + t, 27)
        assert np.array_equal(np.sort(perm), np.arange(n))
        assert np.array_equal(got, labels[perm])
        assert np.array_equal(np.bincount(got, minlength=7), np.bincount(labels, minlength=7))


@pytest.mark.parametrize("C,contract_impl", [(4, None), (9, 0), (9, 2)])
def test_label_permutation_null_matches_the_oracle(dev, C, contract_impl):
    import laris_b200 as la
    from laris_b200.synth import make_dataset
    from oracle import restate as R
    a, lr, c = make_dataset(1, n=3000, G=200, G_u=100, P=40, C=C)
    k, T = 8, 12
    ad = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=k)
    got = la.tl.labelPermutationNull(ad, a, number_nearest_neighbors=k, n_permutations=T, random_seed=5,
                                     contract_impl=contract_impl)
    assert got.shape == (40, C * C) and list(got.index) == list(ad.var_names) and "laris_label_null" in ad.uns
    S = sp.csr_matrix(ad.X).astype(np.float64)
    labels = a.obs["CellTypes"].cat.codes.to_numpy()
    ind, dist = R.knn_graph(a.obsm["X_spatial"], k, False)
    cos0, ge, p = R.label_permutation_null(S, labels, ind, R.decay_weights(dist), C, T, 5)
    g = got.to_numpy()
    assert ((g > 0) & (g <= 1)).all() and np.allclose(g * (T + 1), np.round(g * (T + 1)))
    # counts are integers: identical except where a permuted cosine ties the observed one to ~1e-6
    assert (np.abs(g - p) > 1e-12).mean() < 0.01
    assert (g < 1.0).any() and (g == 1.0).any()                    # the null discriminates
    # runLARIS(label_permutations=...) attaches the same numbers to the long table
    _, tab = la.tl.runLARIS(ad, a, n_nearest_neighbors=k, number_nearest_neighbors=k, n_permutations=20, n_top_lr=40,
                            n_cells_expressed_threshold=10, rng="philox", random_seed=5, label_permutations=T,
                            contract_impl=contract_impl, verbose=False)
    assert "p_value_label_null" in tab.columns
    key = tab["sender"].astype(str) + "::" + tab["receiver"].astype(str)
    want = np.array([got.loc[nm, kk] for nm, kk in zip(tab["interaction_name"].iloc[:200], key.iloc[:200])])
    assert np.array_equal(tab["p_value_label_null"].to_numpy()[:200], want)


# --------------------------------------------------------------------------- K4 over packed score rows
def _sparse_case(n, P, k, R, density, seed=0, dense_row=True):
    from laris_b200 import _kernels as K, engine
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(seed)
    xy = torch.rand((n, 2), dtype=torch.float64, device=dev, generator=g) * 10.0 * np.sqrt(n)
    graph = engine.build_graph(xy, k, False, 100.0)
    ld = K.padded_ld(P)
    S = torch.rand((n, ld), dtype=torch.float32, device=dev, generator=g)
    S *= (torch.rand((n, ld), device=dev, generator=g) < density)
    S[:, P:] = 0
    if dense_row:
        S[n // 2, :P] = 1.5                         # a row longer than one segment
        S[n // 3, :P] = 0                           # and an empty one
    cols, wts, _ = engine.null_graphs(graph, R, 27, "philox") if R else (None, None, None)
    return S, graph, cols, wts


@pytest.mark.parametrize("n,P,k,R,density", [(5000, 37, 3, 1, 0.3), (20000, 100, 5, 2, 0.2), (30000, 700, 10, 3, 0.1),
                                             (8000, 1500, 12, 5, 0.05), (6000, 2500, 40, 2, 0.04), (4001, 2700, 7, 0, 0.1)])
def test_packed_row_statistics_equal_the_dense_passes(n, P, k, R, density):
    """K4 from packed rows (laris_scores_pack + laris_spatial_stats_sparse) against the dense passes (tiles / L2 blocks /
    streaming): same fma chain per element, so sums agree to fp32-partial rounding; counts exactly.  Covers short- and
    long-row variants, 1-4 column quads per owner, k > 32, a dense and an empty row, several launches (R + 1 > 4)."""
    from laris_b200 import _kernels as K, engine
    S, graph, cols, wts = _sparse_case(n, P, k, R, density)
    sm = engine.ScoreMatrix(S, P)
    engine.queue_row_pointers(sm)
    n_entries = int(sm.cache["nnz_get"]()[0])
    sp_ptr = sm.cache["sp_ptr"]
    entries = K.scores_pack(S, P, sp_ptr, n_entries)
    # the packed rows: whole chunks, exactly the non-zeros of S, zero-valued padding behind the last column
    ptr = sp_ptr.cpu().numpy()
    e = entries.cpu().numpy()
    col, val = (e & 0xffffffff).astype(np.int64), ((e >> 32) & 0xffffffff).astype(np.uint32).view(np.float32)
    assert (np.diff(ptr) % 32 == 0).all() and ptr[-1] == n_entries
    real = col < S.shape[1]
    assert (val[~real] == 0).all() and (col[~real] < S.shape[1] + 32).all()
    rows = np.repeat(np.arange(n), np.diff(ptr))
    Sh = S.cpu().numpy()
    assert int(real.sum()) == int((Sh != 0).sum())
    assert np.array_equal(Sh[rows[real], col[real]], val[real])
    back = np.zeros_like(Sh)
    back[rows[real], col[real]] = val[real]
    assert np.array_equal(back, Sh)                                   # every non-zero exactly once
    got = K.spatial_stats_sparse(S, P, sp_ptr, entries, graph.idx, graph.w, graph.order, cols, wts)
    ref = K.spatial_stats_batched(S, P, graph.idx, graph.w, graph.order, cols, wts)
    assert got.shape == ref.shape == (1 + R, 5, P)
    assert torch.equal(got[:, 4], ref[:, 4])
    assert torch.allclose(got, ref, rtol=2e-6, atol=1e-300)
    again = K.spatial_stats_sparse(S, P, sp_ptr, entries, graph.idx, graph.w, graph.order, cols, wts)
    assert torch.equal(got, again)                                    # bitwise reproducible
    if R:
        only_null = K.spatial_stats_sparse(S, P, sp_ptr, entries, None, None, None, cols, wts)
        assert torch.allclose(only_null, ref[1:], rtol=2e-6, atol=1e-300)


def test_whole_path_same_with_packed_and_dense_statistics(monkeypatch):
    """runLARIS with K4 forced through the packed-row pass and through the dense passes: same ranking, same tables."""
    import laris_b200 as la
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(1)
    outs = []
    for mode in ("1", "0"):
        monkeypatch.setenv("LARIS_B200_K4_SPARSE", mode)
        ad = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
        laris_lr, table = la.tl.runLARIS(ad, a, n_nearest_neighbors=c.k, n_repeats=c.R, n_top_lr=c.P,
                                         number_nearest_neighbors=c.k, n_permutations=200, verbose=False, rng="philox")
        outs.append((ad.var.copy(), laris_lr, table))
    (v1, l1, t1), (v0, l0, t0) = outs
    for colname in ("LRSS_Target", "LRSS_Random", "LR_SpatialSpecificity", "total_counts"):
        assert np.allclose(v1[colname].to_numpy(), v0[colname].to_numpy(), rtol=1e-6, atol=1e-12)
    assert np.array_equal(v1["n_cells_by_counts"].to_numpy(), v0["n_cells_by_counts"].to_numpy())
    assert list(l1.index) == list(l0.index)
    assert np.allclose(t1["interaction_score"].to_numpy(), t0["interaction_score"].to_numpy(), rtol=1e-5, atol=1e-12)


def test_sparse_statistics_cost_model():
    from laris_b200 import engine
    # config 3: wide and ~10 % non-zeros -> packed rows; an 8-way shard of it, or 20 % non-zeros -> dense passes
    assert engine.sparse_stats_pays(1_000_000, 2000, 10, 1, 3, 216_000_000)
    assert not engine.sparse_stats_pays(1_000_000, 250, 10, 1, 3, 34_000_000)
    assert not engine.sparse_stats_pays(1_000_000, 2000, 10, 1, 3, 416_000_000)
    assert not engine.sparse_stats_pays(1_000_000, 2000, 40, 1, 3, 216_000_000)


def test_speculative_entry_array_is_validated(monkeypatch):
    """runLARIS sizes the packed-entry array from the previous score matrix of the same shape (no host wait).  When the
    next matrix is denser than that estimate allows, the statistics are redone with the exact size: results equal the
    dense passes either way."""
    import laris_b200 as la
    from laris_b200 import engine
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(1)
    kw = dict(n_nearest_neighbors=c.k, n_repeats=c.R, n_top_lr=c.P, number_nearest_neighbors=c.k, by_celltype=False,
              verbose=False, rng="philox")
    monkeypatch.setenv("LARIS_B200_K4_SPARSE", "0")
    ad0 = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    ref = la.tl.runLARIS(ad0, **kw)
    monkeypatch.setenv("LARIS_B200_K4_SPARSE", "1")
    engine._ENTRY_HINT.clear()
    ad1 = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    first = la.tl.runLARIS(ad1, **kw)                                    # no estimate yet: exact size
    key = (ad1.shape[0], ad1.shape[1])
    exact = engine._ENTRY_HINT[key]
    ad2 = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    second = la.tl.runLARIS(ad2, **kw)                                   # estimate fits
    engine._ENTRY_HINT[key] = exact // 3                                 # pretend the previous matrix was much sparser
    ad3 = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    third = la.tl.runLARIS(ad3, **kw)                                    # estimate too small -> redone
    assert engine._ENTRY_HINT[key] == exact
    for got in (first, second, third):
        assert list(got.index) == list(ref.index)
        assert np.allclose(got["score"].to_numpy(), ref["score"].to_numpy(), rtol=1e-6, atol=1e-12)


# ------------------------------------------------- result materialisation ----
@pytest.mark.parametrize("npair,C,frac", [(7, 3, 0.5), (60, 5, 0.3), (300, 12, 0.05), (2, 1, 1.0), (40, 9, 0.0)])
def test_result_row_order_equals_the_host_mirror(dev, npair, C, frac):
    """T6a/T6b: the device's table order (two stable radix sorts) is exactly the numpy mirror's - ties in both keys,
    thresholded-away rows in melt order - and the gathered columns / codes are the host gathers of that order."""
    from laris_b200 import _kernels as K
    from laris_b200.tools import _utils as U
    rng = np.random.default_rng(npair * 31 + C)
    L = C * C
    inter = np.round(rng.random((npair, L)), 2)                         # two decimals: plenty of ties in the first key
    flat = np.round(inter * rng.random((npair, L)), 1) * (rng.random((npair, L)) < frac)      # ties and zeros in the second
    cols = [rng.random((npair, L)) for _ in range(4)]
    inter_d, flat_d = _d(inter, torch.float64), _d(flat, torch.float64)
    order_d, exotic_d = K.result_row_order(inter_d, flat_d, C)
    want = U._rows_in_table_order(inter, flat, npair, C)
    assert int(exotic_d.item()) == 0
    assert np.array_equal(order_d.cpu().numpy(), want)
    got = K.result_rows_gather(order_d, C, flat_d, *[_d(c, torch.float64) for c in cols])
    assert np.array_equal(got[0].cpu().numpy(), flat.reshape(-1)[want])
    for g, c in zip(got[1:5], cols):
        assert np.array_equal(g.cpu().numpy(), c.reshape(-1)[want])
    pair, rem = np.divmod(want, L)
    assert np.array_equal(got[5].cpu().numpy(), pair)
    assert np.array_equal(got[6].cpu().numpy(), rem // C) and np.array_equal(got[7].cpu().numpy(), rem % C)
    only = K.result_rows_gather(order_d, C, flat_d)                     # no p-value columns
    assert only[1] is None and only[4] is None and np.array_equal(only[0].cpu().numpy(), flat.reshape(-1)[want])
    for bad in (-0.5, np.nan):                                          # scores the device pass hands back to the host
        f2 = flat.copy()
        f2[npair // 2, L // 2] = bad
        assert int(K.result_row_order(inter_d, _d(f2, torch.float64), C)[1].item()) == 1


def test_long_table_same_through_device_and_host_order(dev, tiny, monkeypatch):
    """The DataFrame built from the device-ordered columns equals the one assembled on the host (order, strings, values)."""
    import laris_b200 as la
    from laris_b200 import _kernels as K
    a, lr, c = tiny
    kw = dict(n_nearest_neighbors=8, n_cells_expressed_threshold=20, number_nearest_neighbors=8, n_permutations=100,
              rng="philox", verbose=False)
    lr_adata = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8)
    _, dev_tab = la.tl.runLARIS(lr_adata, a, **kw)
    real = K.result_row_order

    def flagged(inter, flat, C):
        order, exotic = real(inter, flat, C)
        return order, torch.ones_like(exotic)                            # pretend the scores were exotic: host path
    monkeypatch.setattr(K, "result_row_order", flagged)
    _, host_tab = la.tl.runLARIS(lr_adata, a, **kw)
    assert list(dev_tab.columns) == list(host_tab.columns) and len(dev_tab) == 120 * c.C * c.C
    for col in ("sender", "receiver", "ligand", "receptor", "interaction_name"):
        assert dev_tab[col].dtype == host_tab[col].dtype and (dev_tab[col].to_numpy() == host_tab[col].to_numpy()).all(), col
    assert np.allclose(dev_tab["interaction_score"], host_tab["interaction_score"], rtol=1e-12, atol=0)
    assert np.array_equal(dev_tab["p_value"].to_numpy(), host_tab["p_value"].to_numpy())
    assert {"n_genes_by_counts", "total_counts", "pct_counts_in_top_50_genes"} <= set(lr_adata.obs.columns)


def test_pending_score_matrix_joins_on_first_access(dev, tiny):
    """prepareLRInteraction hands the CSR over while its arrays are still coming down; runLARIS scores the resident copy
    without touching them; the first reader waits and sees the oracle's matrix; copies and pickles are plain matrices."""
    import pickle
    import laris_b200 as la
    from laris_b200 import engine
    from oracle import restate as R
    a, lr, _ = tiny
    lr_adata = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8)
    X = lr_adata.X
    assert isinstance(X, engine.PendingCSR) and sp.issparse(X) and X._laris_pending() and X.shape == (1500, 120)
    out = la.tl.runLARIS(lr_adata, None, n_nearest_neighbors=8, by_celltype=False, n_cells_expressed_threshold=20,
                         rng="philox", verbose=False)
    assert X._laris_pending() and len(out) == 120                        # nothing on that path looked at the host arrays
    ind, dist = R.knn_graph(a.obsm["X_spatial"], 8, True)
    S = R.lr_scores(a.X, ind, R.decay_weights(dist), R.gene_index(a.var_names, lr["ligand"]),
                    R.gene_index(a.var_names, lr["receptor"]))
    assert X.nnz == S.nnz and not X._laris_pending()
    assert np.array_equal(X.indptr, S.indptr) and np.array_equal(X.indices, S.indices)
    assert np.allclose(X.data, S.data, rtol=RTOL, atol=0) and X.dtype == np.float32 and X.indptr.dtype == np.int32
    for Y in (X.copy(), pickle.loads(pickle.dumps(X))):
        assert (Y != X).nnz == 0
    assert type(pickle.loads(pickle.dumps(X))) is sp.csr_matrix
    X64 = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8, dtype=np.float64).X
    assert X64.dtype == np.float64 and X64.data.dtype == np.float64 and np.array_equal(X64.data, X.data.astype(np.float64))


def test_row_qc_reaches_into_the_negative_values(dev):
    """Top-N sums when N exceeds the positives plus the zeros of a row (the select then runs below the key of 0.0)."""
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(5)
    n, P = 64, 200
    S = np.zeros((n, K.padded_ld(P)), np.float32)
    S[:, :P] = rng.normal(0.0, 1.0, (n, P)).astype(np.float32) * (rng.random((n, P)) < 0.8)
    S[3, :P] = -np.abs(S[3, :P]) - 1.0                                  # no positives, no zeros
    S[4, :P] = 0.0
    S[4, :5] = [-1.0, -2.0, 3.0, -0.5, 7.0]
    tops = [1, 120, 150, 199]
    total, nnz, top = K.row_qc(torch.from_numpy(S).to(dev), P, tops)
    ref = S[:, :P].astype(np.float64)
    srt = -np.sort(-ref, axis=1)
    assert np.allclose(total.cpu().numpy(), ref.sum(1), rtol=1e-12, atol=1e-12)
    assert np.array_equal(nnz.cpu().numpy(), (ref != 0).sum(1))
    for j, N in enumerate(tops):
        assert np.allclose(top.cpu().numpy()[:, j], srt[:, :N].sum(1), rtol=1e-12, atol=1e-12), N
    view = torch.from_numpy(S).to(dev)[:, 1:]                           # ld not a multiple of 4 / unaligned rows: scalar loads
    Sv = view.contiguous()
    t2 = K.row_qc(Sv, P - 1, [50])[2].cpu().numpy()[:, 0]
    assert np.allclose(t2, (-np.sort(-ref[:, 1:], axis=1))[:, :50].sum(1), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("n,P,k,C,density", [(5000, 37, 3, 4, 0.3), (30000, 700, 10, 40, 0.1), (6000, 2500, 5, 150, 0.04),
                                             (777, 3, 2, 1, 0.5)])
def test_packed_row_celltype_counts_equal_the_dense_pass(n, P, k, C, density):
    """K5b from the packed rows of S: the integer table of the dense pass, exactly - types without cells, labels outside
    [0, C) (skipped by both), a dense and an empty row; rows beyond an under-sized entry array are left out, not misread."""
    from laris_b200 import _kernels as K, engine
    S, _, _, _ = _sparse_case(n, P, k, 0, density)
    sm = engine.ScoreMatrix(S, P)
    engine.queue_row_pointers(sm)
    n_entries = int(sm.cache["nnz_get"]()[0])
    sp_ptr = sm.cache["sp_ptr"]
    entries = K.scores_pack(S, P, sp_ptr, n_entries)
    rng = np.random.default_rng(n + C)
    labels = rng.integers(0, C, n).astype(np.int32)
    if C > 3:
        labels[labels == 2] = 3                                        # a type without cells
        labels[:7] = -1                                                # unlabelled cells
        labels[7:9] = C
    lab_d = torch.from_numpy(labels).cuda()
    dense = K.celltype_nonzero_count(S, P, lab_d, C)
    packed = K.celltype_nonzero_count_packed(sp_ptr, entries, P, lab_d, C)
    assert torch.equal(dense, packed)
    Sh = S.cpu().numpy()[:, :P] != 0
    ok = (labels >= 0) & (labels < C)
    ref = np.zeros((C, P), np.int64)
    np.add.at(ref, labels[ok], Sh[ok].astype(np.int64))
    assert np.array_equal(packed.cpu().numpy(), ref)
    short = K.celltype_nonzero_count_packed(sp_ptr, entries[: n_entries // 2 // 32 * 32].contiguous(), P, lab_d, C)
    assert (short <= packed).all() and (n_entries < 64 or short.sum() < packed.sum())


def test_large_pageable_arrays_go_up_through_the_staging_ring(dev):
    """K.to_device of a pageable array above the 64 MB staging limit (the expression matrix of an ordinary AnnData): the ring
    of page-locked buffers is refilled more than once, the last piece is ragged, dtype conversions happen on the device."""
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(0)
    a = rng.random((230 << 20) // 4 + 12345, dtype=np.float32)          # 8 pieces of 32 MB and a bit
    assert K.PAGEABLE_PIPELINE and not torch.from_numpy(a).is_pinned()
    d = K.to_device(a, torch.float32)
    assert d.dtype == torch.float32 and d.shape == a.shape and np.array_equal(d.cpu().numpy(), a)
    again = K.to_device(a[::-1].copy(), torch.float32)                  # the ring is reused by the next call
    assert np.array_equal(again.cpu().numpy(), a[::-1])
    i = (a[: (70 << 20) // 4] * 1e6).astype(np.int32)
    d64 = K.to_device(i, torch.int64)
    assert d64.dtype == torch.int64 and np.array_equal(d64.cpu().numpy(), i.astype(np.int64))
    m = a[: 3 * 7_000_000].reshape(-1, 3).astype(np.float64)            # 2-D, 168 MB
    dm = K.to_device(m, torch.float64)
    assert dm.shape == m.shape and np.array_equal(dm.cpu().numpy(), m)
