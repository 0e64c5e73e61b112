"""GPU tier: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs and against the committed reference fixtures.

Bar: bit-exact for neighbour indices, distances, CSR structure, permutation
counts; rtol 1e-5 for floating-point scores / statistics / p-values
(BASELINE.json north_star)."""
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
    import laris_b200._lib  # noqa: F401  (fails loudly if the library is missing)
    return torch.device("cuda:0")


def _coords(kind, n, seed=0, dim=2):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((n, dim)) * 10.0 * np.sqrt(n)
    if kind == "hex":
        side = int(np.ceil(np.sqrt(n)))
        ii, jj = np.divmod(np.arange(n), side)
        return np.stack([(jj + 0.5 * (ii & 1)) * 100.0, ii * 86.60254037844386], 1) + rng.normal(0, 1, (n, 2))
    if kind == "bins":
        side = int(np.ceil(np.sqrt(n)))
        ii, jj = np.divmod(np.arange(n), side)
        return np.stack([jj, ii], 1).astype(np.float64) + rng.uniform(-0.01, 0.01, (n, 2))
    if kind == "clustered":            # strongly non-uniform density + far outliers
        c = rng.normal(0, 1, (n, dim)) * rng.choice([0.01, 1.0, 50.0], (n, 1))
        c[:5] += 1e4
        return c
    if kind == "line":                 # degenerate extent in y
        return np.stack([rng.random(n) * 1000.0, np.zeros(n)], 1)
    raise ValueError(kind)


# ------------------------------------------------------------------ K1 ----
@pytest.mark.parametrize("kind,n,dim", [("uniform", 20000, 2), ("hex", 10000, 2), ("bins", 10000, 2),
                                        ("clustered", 6000, 2), ("uniform", 8000, 3), ("line", 3000, 2),
                                        ("uniform", 12, 2)])
@pytest.mark.parametrize("k,include_self", [(10, True), (10, False), (1, False), (31, True)])
def test_knn_bit_exact(dev, kind, n, dim, k, include_self):
    from laris_b200 import engine
    from oracle import restate as R
    if k + (0 if include_self else 1) > n:
        pytest.skip("k too large for n")
    xy = _coords(kind, n, seed=n + k, dim=dim)
    g = engine.build_graph(xy, k, include_self, sigma=100.0)
    ind, dist = R.knn_graph(xy, k, include_self)
    got_i, got_d = g.idx.cpu().numpy(), g.dist.cpu().numpy()
    if k + (0 if include_self else 1) >= n // 2:
        # sklearn switches to brute force here (GEMM-expanded |x|^2+|y|^2-2xy distances,
        # sklearn/neighbors/_base.py _fit): only the KD-tree regime is bit-reproducible
        assert np.array_equal(got_i, ind) and np.allclose(got_d, dist, rtol=1e-9, atol=1e-9)
        return
    assert np.array_equal(got_d, dist), "distances must be bit-identical to sklearn's"
    same = got_i == ind
    if not same.all():                  # only exact distance ties may be ordered differently
        rows = np.unique(np.nonzero(~same)[0])
        for r in rows:
            assert np.array_equal(np.sort(got_i[r]), np.sort(ind[r])) or len(np.unique(dist[r])) < k, r
        assert len(rows) <= max(1, n // 1000)
    order = g.order.cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(n))


def test_knn_matches_reference_fixture(dev, golden):
    from laris_b200 import engine
    g = engine.build_graph(golden["xy"], 8, True)
    assert np.array_equal(g.idx.cpu().numpy(), golden["knn_self_idx"])
    assert np.array_equal(g.dist.cpu().numpy(), golden["knn_self_dist"])
    g2 = engine.build_graph(golden["xy"], 8, False, sigma=100.0)
    assert np.array_equal(g2.idx.cpu().numpy(), golden["knn_noself_idx"])
    assert np.allclose(g2.w64.cpu().numpy(), golden["knn_noself_w"], rtol=1e-12)
    A = g2.to_scipy()
    assert A.shape == (1500, 1500) and A.dtype == np.float64 and (np.diff(A.indptr) == 8).all()


def test_knn_argument_errors(dev):
    from laris_b200 import engine
    from laris_b200._lib import LarisError
    xy = _coords("uniform", 50)
    with pytest.raises(LarisError, match="k out of range"):
        engine.build_graph(xy, 64, False)
    with pytest.raises(LarisError, match="need k"):
        engine.build_graph(xy[:5], 10, True)
    with pytest.raises(ValueError):
        engine.build_graph(np.zeros((10, 4)), 3, True)


def test_decay_weights(dev):
    from laris_b200 import engine
    from oracle import restate as R
    xy = _coords("uniform", 5000, 3)
    g = engine.build_graph(xy, 10, True)
    _, dist = R.knn_graph(xy, 10, True)
    assert np.allclose(g.w64.cpu().numpy(), R.decay_weights(dist), rtol=1e-12)
    assert np.allclose(g.scale.item(), dist.mean() / 2, rtol=1e-13)
    assert np.allclose(g.w.cpu().numpy(), R.decay_weights(dist), rtol=1e-6)
    assert (g.w64.cpu().numpy()[:, 0] == 1.0).all()          # self weight exp(0)


@pytest.mark.parametrize("n", [1, 5, 2047, 2048, 2049, 300001, 5_000_003])
def test_exclusive_scan(dev, n):
    from laris_b200 import _kernels as K
    v = np.random.default_rng(n).integers(0, 1000, n)
    out = K.exclusive_scan(torch.from_numpy(v).to(dev)).cpu().numpy()
    assert out[0] == 0 and np.array_equal(out[1:], np.cumsum(v))


# --------------------------------------------------------------- K2+K3 ----
LAYOUTS = ("dense", "sparse")


def _prepare_both(a, lr, k, layout="auto"):
    import laris_b200 as la
    from oracle import restate as R
    out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=k, layout=layout)
    li = R.gene_index(a.var_names, lr["ligand"])
    ri = R.gene_index(a.var_names, lr["receptor"])
    ind, dist = R.knn_graph(a.obsm["X_spatial"], k, True)
    S = R.lr_scores(a.X, ind, R.decay_weights(dist), li, ri)
    X = sp.csr_matrix(out.X)
    X.sort_indices()
    return out, X, S


def _assert_csr_equal(X, S):
    assert X.shape == S.shape
    assert np.array_equal(X.indptr, S.indptr), "CSR row pointers differ"
    assert np.array_equal(X.indices, S.indices), "CSR column indices differ"
    assert np.allclose(X.data, S.data, rtol=RTOL, atol=0)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_prepare_matches_reference_fixture(dev, golden, tiny, layout):
    import laris_b200 as la
    a, lr, _ = tiny
    out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8, layout=layout)
    X = sp.csr_matrix(out.X)
    X.sort_indices()
    assert np.array_equal(X.indptr, golden["S_indptr"]) and np.array_equal(X.indices, golden["S_indices"])
    assert np.allclose(X.data, golden["S_data"], rtol=RTOL, atol=0)
    assert list(out.var_names[:2]) == [f"{l}::{r}" for l, r in zip(lr["ligand"][:2], lr["receptor"][:2])]
    assert out.obs.equals(a.obs) and np.array_equal(out.obsm["X_spatial"], a.obsm["X_spatial"])
    assert out.obsm["X_spatial"] is not a.obsm["X_spatial"]


@pytest.mark.parametrize("cfg,n,over", [
    (1, 10000, {}),                                              # BASELINE configs[0], full size
    (2, 20000, {}),
    (3, 6000, dict(G=2500, G_u=1500, P=1001)),                   # several column passes, P not a multiple of 8, smem pair table
    (3, 3000, dict(G=4000, G_u=3500, P=300)),                    # wide panel
    (1, 777, dict(G=64, G_u=40, P=7)),                           # ragged group tail, tiny P
    (2, 5001, dict(G=300, G_u=200, P=130)),                      # two 128-pair groups in registers
])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_prepare_vs_oracle(dev, cfg, n, over, layout):
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(cfg, n=n, **over)
    _, X, S = _prepare_both(a, lr, c.k, layout)
    _assert_csr_equal(X, S)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_prepare_signed_expression(dev, layout):
    """Centred (negative) expression values: the expression flag cannot ride in the sign bit."""
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(1, n=3000, G=200, G_u=150, P=90)
    X = a.X.tocsr().astype(np.float32)
    X.data -= 0.7
    a.X = X
    _, Xg, S = _prepare_both(a, lr, c.k, layout)
    assert (S.data < 0).any()
    assert Xg.shape == S.shape and np.array_equal(Xg.indptr, S.indptr) and np.array_equal(Xg.indices, S.indices)
    # products of cancelling sums: compare on the scale of the row, not per element
    assert np.allclose(Xg.data, S.data, rtol=1e-4, atol=1e-5 * np.abs(S.data).max())


def test_prepare_auto_layout(dev):
    from laris_b200 import engine
    from laris_b200.synth import make_dataset
    a, _, _ = make_dataset(2, n=2000)
    assert engine.choose_layout(a.X) == "dense"
    thin = sp.random(2000, 500, density=0.01, format="csr", dtype=np.float32, random_state=1)
    assert engine.choose_layout(thin) == "sparse"
    with pytest.raises(ValueError):
        engine.choose_layout(thin, "blocked")


def test_prepare_edge_cases(dev):
    """Empty cells, unexpressed genes, ligand == receptor, duplicate pairs, k=2, k=33, explicit zeros, float64 X.
    (k=1 is degenerate in the reference itself: mean(d)=0 makes every weight 0/0 = NaN.)"""
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(1, n=1200, G=80, G_u=50, P=30)
    X = sp.lil_matrix(a.X.astype(np.float64))
    X[5:40, :] = 0                     # cells with no expression at all
    X[:, 3] = 0                        # a gene nobody expresses
    X = X.tocsr()
    X.data[::17] = 0.0                 # explicitly stored zeros must not count as expressed
    a.X = X
    lr = pd.concat([lr, pd.DataFrame({"ligand": ["g00003", "g00007", lr["ligand"][0]],
                                      "receptor": ["g00004", "g00007", lr["receptor"][0]]})], ignore_index=True)
    for layout in LAYOUTS:
        for k in (2, 10, 16, 33, 40):
            _, Xg, S = _prepare_both(a, lr, k, layout)
            _assert_csr_equal(Xg, S)
        assert Xg[5:40].nnz == 0


def test_prepare_requires_sparse_and_known_genes(dev, tiny):
    import laris_b200 as la
    a, lr, _ = tiny
    bad = lr.copy()
    bad.loc[0, "ligand"] = "not_a_gene"
    with pytest.raises(KeyError, match="not_a_gene"):
        la.tl.prepareLRInteraction(a, bad)
    d = a.copy()
    d._X = np.asarray(a.X.todense())
    with pytest.raises(TypeError, match="sparse"):
        la.tl.prepareLRInteraction(d, lr)
    with pytest.raises(KeyError, match="X_nope"):
        la.tl.prepareLRInteraction(a, lr, use_rep_spatial="X_nope")


# ------------------------------------------------------------------ K4 ----
def test_spatial_stats_and_nulls(dev, golden, tiny):
    from laris_b200 import engine
    from oracle import restate as R
    n, P = 1500, 120
    S = sp.csr_matrix((golden["S_data"], golden["S_indices"], golden["S_indptr"]), shape=(n, P))
    sm = engine.scores_from_csr(S)
    g = engine.build_graph(golden["xy"], 8, False, sigma=100.0)
    st = engine.observed_stats(sm, g)
    assert np.allclose(engine._cosine(st), golden["gsp"], rtol=RTOL, atol=1e-9)
    assert np.array_equal(st[4].astype(np.int64), golden["nnz"])
    assert np.allclose(st[3], np.asarray(S.sum(0)).ravel(), rtol=1e-6)
    assert np.allclose(st[1], np.asarray(S.multiply(S).sum(0)).ravel(), rtol=1e-6)
    # numpy-stream null == the reference's own random background
    cos, state = engine.null_cosines(sm, g, 3, 27, "numpy")
    assert np.allclose(cos, golden["rand_each"], rtol=RTOL, atol=1e-9)
    assert state is not None
    # philox null == oracle fed the CPU twin's draws
    cos_p, _ = engine.null_cosines(sm, g, 3, 27, "philox")
    ind, dist = R.knn_graph(golden["xy"], 8, False)
    w = R.decay_weights(dist, 100.0)
    ref = [R.null_cosine(S, 8, c, wp.astype(np.float32).astype(np.float64))
           for c, wp in R.null_graphs(n, 8, w.astype(np.float32), 3, 27, "philox")]
    assert np.allclose(cos_p, ref, rtol=RTOL, atol=1e-9)


def test_null_graph_philox_bit_exact(dev):
    from laris_b200 import _kernels as K
    from oracle import rng as orng
    n, k = 30011, 7
    w = torch.rand((n, k), device=dev)
    for rep in (0, 2):
        cols, wout = K.null_graph_philox(n, k, w, rep, 27)
        assert np.array_equal(cols.cpu().numpy().ravel(), orng.philox_null_cols(n, n * k, rep, 27))
        perm = orng.feistel_permutation(n * k, rep, 27)
        assert np.array_equal(wout.cpu().numpy().ravel(), w.cpu().numpy().ravel()[perm])


def test_stats_column_tail_and_large_k(dev):
    """P not a multiple of 4/128, k not a multiple of 4, order=None vs spatial order."""
    from laris_b200 import _kernels as K, engine
    from oracle import restate as R
    rng = np.random.default_rng(5)
    n, P, k = 3001, 131, 13
    S = sp.random(n, P, 0.2, random_state=3, format="csr", dtype=np.float32)
    xy = _coords("uniform", n, 9)
    sm = engine.scores_from_csr(S)
    g = engine.build_graph(xy, k, False, sigma=50.0)
    a = K.spatial_stats(sm.S, P, g.idx, g.w, False, g.order).cpu().numpy()
    b = K.spatial_stats(sm.S, P, g.idx, g.w, False, None).cpu().numpy()
    assert np.allclose(a, b, rtol=1e-12)                         # order only changes fp64 summation order
    ind, dist = R.knn_graph(xy, k, False)
    ref = R.observed_cosine(S.astype(np.float64), ind, R.decay_weights(dist, 50.0))
    assert np.allclose(engine._cosine(a), ref, rtol=RTOL, atol=1e-9)
    del rng


# ------------------------------------------------------------------ K5 ----
def test_celltype_contractions(dev, golden, tiny):
    from laris_b200 import _kernels as K, engine
    from oracle import restate as R
    a, _, c = tiny
    n, P, C = 1500, 120, c.C
    labels = golden["labels"].astype(np.int32)
    S = sp.csr_matrix((golden["S_data"], golden["S_indices"], golden["S_indptr"]), shape=(n, P))
    sm = engine.scores_from_csr(S)
    lab_d = torch.from_numpy(labels).to(dev)
    n_c = np.bincount(labels, minlength=C).astype(np.float64)
    assert np.allclose(engine.celltype_fraction(sm, lab_d, C, n_c), R.celltype_fraction(S, labels, C), rtol=1e-12)
    g = engine.build_graph(golden["xy"], 8, False, None)
    ind, dist = R.knn_graph(golden["xy"], 8, False)
    N_ref = R.neighborhood_composition(labels, ind, R.decay_weights(dist), C)
    N = K.neighborhood_composition(lab_d, g.idx, g.w, C).cpu().numpy()
    assert np.allclose(N.T, N_ref, rtol=1e-5, atol=1e-7)
    for impl in (None, 2, 0, 1):                    # default (= key-gather here) / key-gather / tcgen05 / fp32 CUDA cores
        raw, cos = engine.neighborhood_specificity(sm, lab_d, C, g, 100.0, impl=impl)
        assert np.allclose(cos, R.celltype_pair_cosine(S, N_ref), rtol=RTOL, atol=1e-8), impl
        # the reference's own table.  raw ~ cos^3: the default path holds the 1e-5 bar; the tensor-core path
        # (bf16 operand split, truncating TMEM accumulation) is documented at 2e-5 on this cubed quantity
        assert np.allclose(raw, golden["ctc_raw"], rtol=2e-5 if impl == 0 else RTOL, atol=1e-10), impl
    # COSG gene scores
    X = sp.csr_matrix(a.X)
    genes = np.unique(np.concatenate([golden["lig_idx"], golden["rec_idx"]]))
    xu = engine.select_genes(X, genes)
    got = engine.cosg_gene_scores(xu, lab_d, C, n_c.astype(np.int64), 100.0, 0.1, True)
    ref = R.cosg_scores(X[:, genes], labels, C, 100.0, 0.1, True)
    assert np.array_equal(got == -1, ref == -1)
    assert np.allclose(got, ref, rtol=RTOL, atol=1e-9)


@pytest.mark.parametrize("n,P,C,perm", [
    (300, 40, 8, False),          # one tile, ragged everything
    (1000, 128, 9, True),         # P exactly one tile; C*C < 128
    (4097, 500, 20, True),        # several row/column tiles, n % 64 != 0
    (9000, 131, 40, False),       # column tiles straddling cell types
    (2500, 70, 64, True),         # largest supported number of cell types
    (3000, 33, 1, False),         # single cell type
])
def test_pair_contract_tensor_cores(dev, n, P, C, perm):
    """tcgen05 contraction (impl=0) against a float64 contraction and against the fp32 CUDA-core kernel."""
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(n + P + C)
    ld = K.padded_ld(P)
    S = np.zeros((n, ld), np.float32)
    S[:, :P] = (rng.random((n, P)) < 0.3) * rng.gamma(2.0, 1.0, (n, P)).astype(np.float32)
    N = np.zeros((n, C), np.float32)
    for _ in range(3):
        N[np.arange(n), rng.integers(0, C, n)] += rng.random(n).astype(np.float32)
    N[::50] = 0.0                                                    # cells with an empty neighbourhood profile
    Sd, Nd = torch.from_numpy(S).to(dev), torch.from_numpy(N).to(dev)
    order = torch.from_numpy(rng.permutation(n).astype(np.int32)).to(dev) if perm else None
    num0, q0 = K.celltype_pair_contract(Sd, P, Nd, 0, order=order)
    num1, q1 = K.celltype_pair_contract(Sd, P, Nd, 1)
    Q = (N[:, :, None].astype(np.float64) * N[:, None, :]).reshape(n, C * C)
    ref = S[:, :P].astype(np.float64).T @ Q
    a0, a1 = num0.cpu().numpy(), num1.cpu().numpy()
    assert np.array_equal(a0 == 0, ref == 0)                         # exact zeros stay exact
    # split-bf16 operands: every product is exact to 2^-17, sums of positive terms inherit that bound
    assert np.allclose(a0, ref, rtol=RTOL, atol=0)
    assert np.allclose(a1, ref, rtol=1e-6, atol=0)
    assert np.allclose(q0.cpu().numpy(), (Q * Q).sum(0), rtol=1e-6)
    if perm:                                                         # the cell order only changes fp32 summation order
        num2, _ = K.celltype_pair_contract(Sd, P, Nd, 0)
        assert np.allclose(num2.cpu().numpy(), a0, rtol=2e-6, atol=0)


def test_pair_contract_argument_errors(dev):
    from laris_b200 import _kernels as K
    from laris_b200._lib import LarisError
    S = torch.zeros((64, 8), device=dev)
    N = torch.zeros((64, 65), device=dev)
    with pytest.raises(LarisError, match="C <= 64"):
        K.celltype_pair_contract(S, 8, N, 0)
    num, _ = K.celltype_pair_contract(S, 8, N)                        # auto: falls to the CUDA-core kernel
    assert float(num.abs().sum()) == 0.0


# --------------------------------------------------------------- K6/K7 ----
def test_pvalue_counts_exact(dev):
    from laris_b200 import engine
    from oracle import restate as R
    from oracle import rng as orng
    rng = np.random.default_rng(11)
    Gn, P, nb, N = 9, 57, 30, 1003
    table = rng.random((Gn, P)) * (rng.random((Gn, P)) < 0.4)
    active = (rng.random(P) < 0.8).astype(np.uint8)
    table[:, active == 0] = 0.0
    bg = np.stack([rng.permutation(P)[:nb] for _ in range(P)]).astype(np.int32)
    # philox: the oracle is fed the CPU twin's draws
    ge, nz, both = engine.pvalue_counts(table, bg, active, N, rng="philox", seed=27)
    rge, rnz, rboth = R.pvalue_counts(table, active.astype(bool), bg,
                                      lambda g, pidx: orng.philox_pvalue_draws(g * P + pidx, N, nb, 27))
    m = active.astype(bool)
    assert np.array_equal(ge[:, m], rge[:, m]) and np.array_equal(nz[:, m], rnz[:, m])
    assert np.array_equal(both[:, m], rboth[:, m])
    assert (ge[:, ~m] == 0).all()
    # host-supplied (numpy-stream) draws
    draws = rng.integers(0, nb, (Gn, P, N)).astype(np.uint8)
    ge2, nz2, both2 = engine.pvalue_counts(table, bg, active, N, rng="numpy", draws=draws)
    rge2, _, rboth2 = R.pvalue_counts(table, m, bg, lambda g, pidx: draws[g][pidx].astype(np.int64))
    assert np.array_equal(ge2[:, m], rge2[:, m]) and np.array_equal(both2[:, m], rboth2[:, m])


def test_pvalue_counts_shard_invariance(dev):
    """Permutation shards and pair shards add up to the unsharded counts (multi-GPU contract)."""
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(3)
    Gn, P, nb, N = 4, 40, 30, 1000
    table = torch.from_numpy(rng.random((Gn, P))).to(dev)
    bg = torch.from_numpy(np.stack([rng.permutation(P)[:nb] for _ in range(P)]).astype(np.int32)).to(dev)
    full = K.pvalue_counts(table, bg, N, seed=5)[0].cpu().numpy()
    acc = None
    for t0, cnt in ((0, 333), (333, 334), (667, 333)):
        acc = K.pvalue_counts(table, bg, cnt, seed=5, t0=t0, out=acc)
    assert np.array_equal(acc[0].cpu().numpy(), full)


def test_bh_fdr(dev):
    from laris_b200 import engine
    from oracle import restate as R
    rng = np.random.default_rng(2)
    for L in (1, 5, 33, 500, 2000):
        p = np.round(rng.random((6, L)) ** 2, 6)
        p[0, : L // 2] = p[0, 0]                      # heavy ties
        sel = rng.random((6, L)) < 0.7
        sel[1] = False
        got = engine.bh_fdr(p, sel.astype(np.uint8))
        for g in range(6):
            exp = np.ones(L)
            if sel[g].any():
                exp[sel[g]] = R.bh_fdr(p[g, sel[g]])
            assert np.allclose(got[g], exp, rtol=1e-12), (L, g)


# ------------------------------------------------------- whole path ----
def _keyed(res, lr_names, cats):
    C = len(cats)
    pid = pd.Index(lr_names).get_indexer(res["interaction_name"].to_numpy())
    key = (pid * C + pd.Index(cats).get_indexer(res["sender"])) * C + pd.Index(cats).get_indexer(res["receiver"])
    return key, np.argsort(key)


def test_runlaris_matches_reference_fixture(dev, golden, tiny):
    """prepare -> runLARIS(rng='numpy') against the unmodified reference's outputs."""
    import laris_b200 as la
    a, lr, c = tiny
    lr_adata = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8)
    np.random.seed(999)
    laris_lr, res = la.tl.runLARIS(lr_adata, a, n_nearest_neighbors=8, random_seed=27, n_repeats=3, mu=1, sigma=100,
                                   n_cells_expressed_threshold=20, n_top_lr=4000, number_nearest_neighbors=8,
                                   mu_celltype=100, n_neighbors_permutation=30, n_permutations=200, rng="numpy",
                                   verbose=False)
    assert np.allclose(lr_adata.var["LRSS_Target"], golden["gsp"], rtol=RTOL, atol=1e-9)
    assert np.allclose(lr_adata.var["LRSS_Random"], golden["rand"], rtol=RTOL, atol=1e-9)
    assert np.array_equal(lr_adata.var["n_cells_by_counts"], golden["nnz"])
    lr_names = np.asarray(lr_adata.var_names)
    assert list(laris_lr.columns) == ["ligand", "receptor", "score", "Rank"]
    assert np.array_equal(laris_lr.index.to_numpy(), lr_names[golden["rank_order"]])     # permutation-derived ranks
    assert np.allclose(laris_lr["score"], golden["rank_score"], rtol=RTOL, atol=1e-8)
    assert list(res.columns) == ["sender", "receiver", "interaction_score", "ligand", "receptor", "interaction_name",
                                 "p_value", "p_value_fdr", "nlog10_p_value_fdr"]
    assert (np.diff(res["interaction_score"].to_numpy()) <= 0).all()
    key, o = _keyed(res, lr_names, list(a.obs["CellTypes"].cat.categories))
    assert np.array_equal(key[o], golden["tbl_key"])
    sc = res["interaction_score"].to_numpy()[o]
    assert np.array_equal(sc == 0, golden["tbl_score"] == 0)
    assert np.allclose(sc, golden["tbl_score"], rtol=RTOL, atol=1e-12)
    assert np.allclose(res["p_value"].to_numpy()[o], golden["tbl_p"], rtol=1e-12), "p-values are count-derived: exact"
    assert np.allclose(res["p_value_fdr"].to_numpy()[o], golden["tbl_fdr"], rtol=1e-9)
    assert np.allclose(res["nlog10_p_value_fdr"].to_numpy()[o], golden["tbl_nlog"], rtol=1e-9, atol=1e-12)
    # background sets (30-NN in (mean, variance) space) equal the reference's KD-tree result
    from laris_b200.tools import _utils as U
    assert np.array_equal(U._prepare_background_interactions(lr_adata, 30).to_numpy(), golden["bg"])
    u = lr_adata.uns["cosg_lr"]
    assert set(u) >= {"params", "COSG", "names", "scores"} and u["params"]["method"] == "COSG"
    assert u["COSG"].shape == (120, 2 * c.C * c.C) and u["scores"].dtype[0] == np.float32
    # the reference leaves np.random advanced exactly like this
    from oracle import rng as orng
    rs = orng.numpy_null_graph(1500, 8, orng.numpy_repeat_seeds(27, 3)[-1])[2]
    for _ in range(c.C * c.C):
        rs.randint(0, 30, size=(120, 200))
    assert np.random.randint(0, 1 << 30) == rs.randint(0, 1 << 30)


def test_runlaris_spatial_only_and_uploaded_scores(dev, golden, tiny):
    """by_celltype=False returns one frame; a caller-built lr_adata (no device cache) gives the same numbers."""
    import laris_b200 as la
    from laris_b200._anndata import AnnDataLite
    a, lr, _ = tiny
    S = sp.csr_matrix((golden["S_data"], golden["S_indices"], golden["S_indptr"]), shape=(1500, 120))
    names = (lr["ligand"] + "::" + lr["receptor"]).to_numpy(dtype=object)
    lr_adata = AnnDataLite(S, obs=a.obs.copy(), var=pd.DataFrame(index=pd.Index(names, dtype=object)),
                           obsm={"X_spatial": a.obsm["X_spatial"]})
    out = la.tl.runLARIS(lr_adata, None, n_nearest_neighbors=8, by_celltype=False, n_cells_expressed_threshold=20,
                         verbose=False)
    assert isinstance(out, pd.DataFrame) and len(out) == 120
    assert np.allclose(lr_adata.var["LRSS_Target"], golden["gsp"], rtol=RTOL, atol=1e-9)
    assert np.array_equal(out.index.to_numpy(), names[golden["rank_order"]])
    out_p = la.tl.runLARIS(lr_adata, None, n_nearest_neighbors=8, by_celltype=False, n_cells_expressed_threshold=20,
                           rng="philox", verbose=False)
    assert len(out_p) == 120 and np.isfinite(out_p["score"]).all()


def test_runlaris_philox_conditional(dev, tiny):
    """Device RNG end to end, conditional p-values, no prefilter: well-formed and reproducible."""
    import laris_b200 as la
    a, lr, c = tiny
    lr_adata = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=8)
    kw = dict(n_nearest_neighbors=8, n_cells_expressed_threshold=20, number_nearest_neighbors=8, n_permutations=300,
              rng="philox", use_conditional_pvalue=True, prefilter_fdr=False, verbose=False)
    _, r1 = la.tl.runLARIS(lr_adata, a, **kw)
    _, r2 = la.tl.runLARIS(lr_adata, a, **kw)
    # reproducible up to fp64 atomic summation order in the two cell-type contractions
    assert r1[["sender", "receiver", "interaction_name"]].equals(r2[["sender", "receiver", "interaction_name"]])
    assert np.allclose(r1["interaction_score"], r2["interaction_score"], rtol=1e-12, atol=0)
    assert np.array_equal(r1["p_value"].to_numpy(), r2["p_value"].to_numpy())
    assert len(r1) == 120 * c.C * c.C
    p = r1["p_value"].to_numpy()
    assert ((p > 0) & (p <= 1)).all() and (p[r1["interaction_score"].to_numpy() == 0] == 1).all()
    assert ((r1["p_value_fdr"] >= r1["p_value"].round(6) - 1e-12) & (r1["p_value_fdr"] <= 1)).all()


def test_pp_helpers(dev):
    import laris_b200 as la
    rng = np.random.default_rng(0)
    A, B = rng.random((40, 300)), rng.random((40, 300))
    A[3] = 0
    ref = np.einsum("ij,ij->i", A, B) / np.maximum(np.linalg.norm(A, axis=1) * np.linalg.norm(B, axis=1), 1e-300)
    ref[3] = 0
    assert np.allclose(la.pp._rowwise_cosine_similarity(A, B), ref, rtol=RTOL)
    assert np.allclose(la.pp._rowwise_cosine_similarity(sp.csr_matrix(A), sp.csr_matrix(B)), ref, rtol=RTOL)
    m = sp.csr_matrix(np.array([[1., 0, 2, 0, 1], [0, 1, 1, 0, 2], [2, 0, 0, 1, 1]]))
    prod, names = la.pp._pairwise_row_multiply(m, ["TypeA", "TypeB", "TypeC"])
    assert prod.shape == (9, 5) and list(names[:3]) == ["TypeA::TypeA", "TypeA::TypeB", "TypeA::TypeC"]
    d = m.toarray()
    assert np.array_equal(prod.toarray(), (d[:, None, :] * d[None, :, :]).reshape(9, 5))


def test_utils_wrappers(dev, golden, tiny):
    from laris_b200.tools import _utils as U
    a, _, _ = tiny
    A = U._build_adjacency_matrix(a, use_rep="X_spatial", n_nearest_neighbors=8, sigma=100)
    assert np.array_equal(A.indices.reshape(1500, 8), golden["knn_noself_idx"])
    S = sp.csr_matrix((golden["S_data"], golden["S_indices"], golden["S_indptr"]), shape=(1500, 120))
    bgl = U._generate_random_background(a, A, S.T.tocsr(), n_nearest_neighbors=8, n_repeats=3, random_seed=27)
    assert np.allclose(np.asarray(bgl), golden["rand_each"], rtol=RTOL, atol=1e-9)
    R_ = U._build_random_adjacency_matrix(a, A, n_nearest_neighbors=8, random_seed=5)
    assert R_.shape == A.shape and np.isclose(R_.sum(), A.sum())


# ------------------------------------------- full-size parity ----
def test_full_size_config2_vs_oracle(dev):
    """BASELINE configs[1] at FULL size (1e5 cells x 500 pairs): prepareLRInteraction against the oracle (exact CSR
    structure, scores to 1e-5), the observed statistics / null cosines against the oracle on the same scores, plus
    the size-independent invariants (checksum of checksums, linearity)."""
    import laris_b200 as la
    from laris_b200 import engine
    from laris_b200.synth import make_dataset
    from laris_b200.tools import _utils as U
    from oracle import restate as R
    a, lr, c = make_dataset(2)
    out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k)
    X = out.X
    assert X.shape == (c.n, c.P) and X.has_sorted_indices and (X.data != 0).all()
    li, ri = R.gene_index(a.var_names, lr["ligand"]), R.gene_index(a.var_names, lr["receptor"])
    ind, dist = R.knn_graph(a.obsm["X_spatial"], c.k, True)
    S = R.lr_scores(a.X, ind, R.decay_weights(dist), li, ri)
    _assert_csr_equal(X, S)
    sm = U._scores_of(out)
    g = engine.build_graph(a.obsm["X_spatial"], c.k, False, sigma=100.0)
    st_all, _ = engine.spatial_stats_all(sm, g, 1, 27, "philox")
    st = st_all[0].cpu().numpy()
    ind2, d2 = R.knn_graph(a.obsm["X_spatial"], c.k, False)
    w2 = R.decay_weights(d2, 100.0)
    assert np.array_equal(g.idx.cpu().numpy(), ind2)
    assert np.allclose(engine._cosine(st), R.observed_cosine(S, ind2, w2), rtol=RTOL, atol=1e-9)
    cols, wp = next(iter(R.null_graphs(c.n, c.k, w2.astype(np.float32), 1, 27, "philox")))
    ref_null = R.null_cosine(S, c.k, cols, wp.astype(np.float32).astype(np.float64))
    assert np.allclose(engine._cosine(st_all[1].cpu().numpy()), ref_null, rtol=RTOL, atol=1e-9)
    # checksum of checksums: device column sums == host sums of the materialised CSR
    assert np.allclose(st[3], np.asarray(X.sum(0)).ravel(), rtol=1e-5)
    assert np.array_equal(st[4].astype(np.int64), np.asarray(X.getnnz(0)).ravel())
    # linearity of the diffusion: scaling X by 2 scales every score by 4
    b = a.copy()
    b.X = a.X * 2.0
    X2 = la.tl.prepareLRInteraction(b, lr, number_nearest_neighbors=c.k).X
    assert np.array_equal(X2.indices, X.indices) and np.allclose(X2.data, 4.0 * X.data, rtol=1e-6)


# ----------------------------------------------------------------- K1r ----
@pytest.mark.parametrize("kind,n,dim,radius", [("uniform", 20000, 2, 18.0), ("hex", 10000, 2, 105.0), ("bins", 10000, 2, 1.05),
                                               ("clustered", 6000, 2, 0.5), ("uniform", 6000, 3, 60.0),
                                               ("line", 3000, 2, 1.0), ("uniform", 40, 2, 1e-3)])
@pytest.mark.parametrize("include_self", [True, False])
def test_radius_graph_bit_exact(dev, kind, n, dim, radius, include_self):
    """K1r against sklearn.radius_neighbors_graph: identical CSR structure and bit-identical distances."""
    from laris_b200 import _kernels as K
    from oracle import restate as R
    xy = _coords(kind, n, seed=n + 7, dim=dim)
    indptr, indices, dist, order = K.radius_graph(K.to_device(xy, torch.float64), radius, include_self)
    r_indptr, r_indices, r_dist = R.radius_graph(xy, radius, include_self)
    assert np.array_equal(indptr.cpu().numpy(), r_indptr)
    assert np.array_equal(indices.cpu().numpy(), r_indices)
    assert np.array_equal(dist.cpu().numpy(), r_dist)
    assert np.array_equal(np.sort(order.cpu().numpy()), np.arange(n))


def test_radius_graph_scores_match_oracle(dev):
    """prepareLRInteraction(radius=...) == W X gather-product-mask with the sklearn radius graph as W."""
    import laris_b200 as la
    from laris_b200 import engine
    from laris_b200.synth import make_dataset
    from oracle import restate as R
    a, lr, c = make_dataset(5, n=6000, G=200, G_u=120, P=90, C=4)        # bin lattice, pitch 1
    radius = 1.5                                                           # 8-neighbourhood + self, ragged at the border
    xy = a.obsm["X_spatial"]
    g = engine.build_radius_graph(xy, radius, True)
    indptr, indices, dist = R.radius_graph(xy, radius, True)
    w = R.decay_weights(dist)                                              # s = mean(stored d)/2, zeros on the diagonal included
    deg = np.diff(indptr)
    assert g.k == deg.max() and deg.min() < deg.max()
    A = g.to_scipy()
    assert np.array_equal(A.indptr, indptr) and np.array_equal(A.indices, indices) and np.allclose(A.data, w, rtol=1e-12)
    # padded rows: own row with weight 0
    gi, gw = g.idx.cpu().numpy(), g.w.cpu().numpy()
    short = np.flatnonzero(deg < deg.max())[:50]
    for i in short:
        assert (gi[i, deg[i]:] == i).all() and (gw[i, deg[i]:] == 0).all()
    out = la.tl.prepareLRInteraction(a, lr, radius=radius)
    W = sp.csr_matrix((w, indices, indptr), shape=(c.n, c.n))
    S = R.lr_scores_csr_graph(a.X, W, R.gene_index(a.var_names, lr["ligand"]), R.gene_index(a.var_names, lr["receptor"]))
    X = sp.csr_matrix(out.X)
    X.sort_indices()
    assert np.array_equal(X.indptr, S.indptr) and np.array_equal(X.indices, S.indices)
    assert np.allclose(X.data, S.data, rtol=RTOL, atol=0)
    # the neighbourhood composition takes the same padded rows
    labels = a.obs["CellTypes"].cat.codes.to_numpy().astype(np.int32)
    from laris_b200 import _kernels as K
    N = K.neighborhood_composition(K.to_device(labels, torch.int32), g.idx, g.w, c.C).cpu().numpy()
    onehot = sp.csr_matrix((np.ones(c.n), (labels, np.arange(c.n))), shape=(c.C, c.n))
    assert np.allclose(N.T, (onehot @ W.T).toarray(), rtol=1e-5, atol=1e-7)


def test_radius_graph_argument_errors(dev):
    from laris_b200 import engine
    from laris_b200._lib import LarisError
    xy = _coords("uniform", 100)
    with pytest.raises(LarisError, match="radius must be positive"):
        engine.build_radius_graph(xy, 0.0, True)
    with pytest.raises(ValueError, match="without a neighbour"):
        engine.build_radius_graph(xy, 1e-9, False)


# ------------------------------------------------ multi-subunit complexes ----
@pytest.mark.parametrize("layout", ["sparse", "dense"])
@pytest.mark.parametrize("rule", ["min", "gmean"])
@pytest.mark.parametrize("signed", [False, True])
def test_complex_aggregation_matches_oracle(dev, layout, rule, signed):
    """In-kernel subunit aggregation (extension) against the numpy oracle; single-gene rows stay the reference's."""
    import laris_b200 as la
    from laris_b200.synth import make_dataset
    from oracle import restate as R
    a, lr, c = make_dataset(2, n=3000, G=120, G_u=80, P=40, C=4, density=0.35)
    if signed:                                                     # negative expression -> bitmap flag path
        X = a.X.copy().astype(np.float32)
        X.data[::7] *= -1.0
        a.X = X
    rng = np.random.default_rng(9)
    names = np.asarray(a.var_names, dtype=object)
    lig, rec, lsets, rsets = list(lr["ligand"]), list(lr["receptor"]), [], []
    col = {nm: i for i, nm in enumerate(names)}
    for p in range(len(lig)):                                      # every third row gets a 2-3 subunit receptor / ligand
        ls, rs = [col[lig[p]]], [col[rec[p]]]
        if p % 3 == 0:
            rs = list(rng.choice(80, int(rng.integers(2, 4)), replace=False))
            rec[p] = "+".join(names[rs])
        if p % 5 == 0:
            ls = list(rng.choice(80, 2, replace=False))
            lig[p] = "+".join(names[ls])
        lsets.append(ls)
        rsets.append(rs)
    lr2 = pd.DataFrame({"ligand": lig, "receptor": rec})
    out = la.tl.prepareLRInteraction(a, lr2, number_nearest_neighbors=8, complex_sep="+", complex_rule=rule, layout=layout)
    ind, dist = R.knn_graph(a.obsm["X_spatial"], 8, True)
    S = R.lr_scores_complex(a.X, R.graph_csr(ind, R.decay_weights(dist), c.n), lsets, rsets, rule)
    X = sp.csr_matrix(out.X)
    X.sort_indices()
    assert list(out.var_names) == [f"{l}::{r}" for l, r in zip(lig, rec)]
    if not signed:
        assert np.array_equal(X.indptr, S.indptr) and np.array_equal(X.indices, S.indices)
        assert np.allclose(X.data, S.data, rtol=RTOL, atol=0)
    else:                                  # mixed signs: sums may cancel to ~0, compare values with an absolute floor
        assert np.allclose(X.toarray(), S.toarray(), rtol=1e-4, atol=1e-5)
    # rows without a complex are exactly the single-gene path
    plain = [p for p in range(len(lig)) if "+" not in lig[p] and "+" not in rec[p]]
    ref = la.tl.prepareLRInteraction(a, lr.iloc[plain].reset_index(drop=True), number_nearest_neighbors=8, layout=layout)
    assert np.array_equal(sp.csr_matrix(out.X)[:, plain].toarray(), sp.csr_matrix(ref.X).toarray())


def test_complex_argument_errors(dev):
    import laris_b200 as la
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(2, n=500, G=60, G_u=40, P=10, C=3)
    bad = lr.copy()
    bad.loc[0, "receptor"] = "g00001+nope"
    with pytest.raises(KeyError):
        la.tl.prepareLRInteraction(a, bad, complex_sep="+")
    with pytest.raises(ValueError, match="complex_rule"):
        la.tl.prepareLRInteraction(a, lr, complex_sep="+", complex_rule="max")


# ------------------------------------------------------------ per-cell QC ----
def test_row_qc_matches_numpy(dev):
    """Per-cell totals, counts and top-N sums (scanpy's obs QC columns) against a sort-based numpy computation."""
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(21)
    n, P = 700, 333
    S = np.zeros((n, K.padded_ld(P)), np.float32)
    S[:, :P] = (rng.random((n, P)) < 0.3) * rng.gamma(2.0, 1.0, (n, P)).astype(np.float32)
    S[5, :P] = 0.0                                        # an empty cell
    S[6, :P] = 1.25                                       # all ties
    S[7, :P] = -S[8, :P]                                  # negative values order correctly too
    tops = [1, 50, 100, 333, 500]
    total, nnz, top = K.row_qc(torch.from_numpy(S).to(dev), P, tops[:4])
    ref = S[:, :P].astype(np.float64)
    assert np.allclose(total.cpu().numpy(), ref.sum(1), rtol=1e-12, atol=1e-12)
    assert np.array_equal(nnz.cpu().numpy(), (ref != 0).sum(1))
    srt = -np.sort(-ref, axis=1)
    for j, N in enumerate(tops[:4]):
        assert np.allclose(top.cpu().numpy()[:, j], srt[:, :min(N, P)].sum(1), rtol=1e-12, atol=1e-12), N
    _, _, top5 = K.row_qc(torch.from_numpy(S).to(dev), P, [500])
    assert np.allclose(top5.cpu().numpy()[:, 0], ref.sum(1), rtol=1e-12, atol=1e-12)        # N >= P: everything


@pytest.mark.parametrize("n,P,C,coherent", [(5000, 131, 7, False), (20000, 500, 40, True), (777, 3, 1, False),
                                            (9001, 64, 150, False)])
def test_celltype_nonzero_count_exact(dev, n, P, C, coherent):
    """K5b: #{cells of type c with S[i,p] != 0} is an integer table - exact for any label arrangement and any C."""
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(n + P)
    S = np.zeros((n, K.padded_ld(P)), np.float32)
    S[:, :P] = (rng.random((n, P)) < 0.25) * rng.random((n, P)).astype(np.float32)
    labels = (np.arange(n) * C // n if coherent else rng.integers(0, C, n)).astype(np.int32)
    got = K.celltype_nonzero_count(torch.from_numpy(S).to(dev), P, torch.from_numpy(labels).to(dev), C).cpu().numpy()
    ref = np.zeros((C, P), np.int64)
    np.add.at(ref, labels, (S[:, :P] != 0).astype(np.int64))
    assert np.array_equal(got, ref)


def test_runlaris_on_radius_graphs(dev):
    """runLARIS(radius=...): observed statistic, random-graph null (numpy stream on the padded weight slots) and the
    neighbourhood co-localisation cosines against the oracle fed the sklearn radius graph."""
    import laris_b200 as la
    from laris_b200 import engine, _kernels as K
    from laris_b200.synth import make_dataset
    from oracle import restate as R, rng as orng
    a, lr, c = make_dataset(5, n=4900, G=150, G_u=90, P=60, C=4, density=0.2)
    radius, n = 1.5, 4900
    lr_adata = la.tl.prepareLRInteraction(a, lr, radius=radius)
    la.tl.runLARIS(lr_adata, a, by_celltype=False, n_repeats=2, random_seed=27, sigma=100.0, n_top_lr=60,
                   n_cells_expressed_threshold=5, rng="numpy", verbose=False, radius=radius)
    S = sp.csr_matrix(lr_adata.X).astype(np.float64)
    indptr, indices, dist = R.radius_graph(a.obsm["X_spatial"], radius, False)
    w = R.decay_weights(dist, 100.0)
    W = sp.csr_matrix((w, indices, indptr), shape=(n, n))
    gsp = R.column_cosine(S, W @ S)
    assert np.allclose(lr_adata.var["LRSS_Target"], gsp, rtol=RTOL, atol=1e-9)
    deg = np.diff(indptr)
    kmax = int(deg.max())
    wpad = np.zeros((n, kmax))
    wpad[np.repeat(np.arange(n), deg), np.concatenate([np.arange(d) for d in deg])] = w
    wpad = wpad.astype(np.float32).astype(np.float64).ravel()           # the device shuffles the fp32 weights
    reps = []
    for s in orng.numpy_repeat_seeds(27, 2):
        cols, perm, _ = orng.numpy_null_graph(n, kmax, s)
        reps.append(R.null_cosine(S, kmax, cols, wpad[perm]))
    assert np.allclose(lr_adata.var["LRSS_Random"], np.mean(reps, axis=0), rtol=RTOL, atol=1e-9)
    # neighbourhood profile on the radius graph with the mean/2 decay
    labels = a.obs["CellTypes"].cat.codes.to_numpy().astype(np.int32)
    w3 = R.decay_weights(dist)
    onehot = sp.csr_matrix((np.ones(n), (labels, np.arange(n))), shape=(c.C, n))
    N_ref = np.asarray((onehot @ sp.csr_matrix((w3, indices, indptr), shape=(n, n)).T).todense())
    from laris_b200.tools import _utils as U
    g3 = engine.build_radius_graph(a.obsm["X_spatial"], radius, False, None)
    _, cos = engine.neighborhood_specificity(U._scores_of(lr_adata), K.to_device(labels, torch.int32), c.C, g3, 100.0)
    assert np.allclose(cos, R.celltype_pair_cosine(S, N_ref), rtol=RTOL, atol=1e-8)
    # and the whole cell-type step runs on it
    laris_lr, table = la.tl.runLARIS(lr_adata, a, n_repeats=2, n_top_lr=60, n_cells_expressed_threshold=5, n_permutations=50,
                                     rng="philox", verbose=False, radius=radius)
    assert len(laris_lr) == 60 and len(table) == 60 * c.C * c.C and np.isfinite(table["interaction_score"]).all()


def test_non_finite_coordinates(dev):
    """Host arrays are rejected up front; a device tensor with a NaN is flagged in the outputs (the build is asynchronous)."""
    from laris_b200 import engine, _kernels as K
    xy = _coords("uniform", 500)
    xy[17, 1] = np.nan
    with pytest.raises(ValueError, match="NaN or infinite"):
        engine.build_graph(xy, 5, True)
    idx, dist, _ = K.knn_graph(torch.from_numpy(xy).to(dev), 5, True)
    assert (idx.cpu().numpy() == -1).all() and np.isnan(dist.cpu().numpy()).all()


@pytest.mark.parametrize("nb", [7, 30, 32, 33, 64, 65, 200])
def test_pvalue_counts_candidate_set_sizes(dev, nb):
    """Every code path of K6 (register masks / flag bytes, 3 or 1 draws per Philox word, ge-only or all three
    counts) against the oracle fed the CPU twin of the device draws."""
    from laris_b200 import engine
    from oracle import restate as R, rng as orng
    rng = np.random.default_rng(nb)
    groups, P, N = 5, max(nb + 1, 40), 157
    table = rng.random((groups, P)) * (rng.random((groups, P)) < 0.6)
    active = np.ones(P, np.uint8)
    bg = np.stack([rng.permutation(P)[:nb] for _ in range(P)]).astype(np.int32)
    ref = R.pvalue_counts(table, active.astype(bool), bg.astype(np.int64),
                          lambda g, pidx: orng.philox_pvalue_draws(g * P + pidx, N, nb, 11))
    ge, nz, both = engine.pvalue_counts(table, bg, active, N, rng="philox", seed=11)
    assert np.array_equal(ge, ref[0]) and np.array_equal(nz, ref[1]) and np.array_equal(both, ref[2])
    ge1, _, _ = engine.pvalue_counts(table, bg, active, N, rng="philox", seed=11, only_ge=True)
    assert np.array_equal(ge1, ref[0])


@pytest.mark.parametrize("layout", LAYOUTS)
def test_captured_scoring_replays_the_step(dev, layout):
    """engine.CapturedScoring: the CUDA-graph replay of the step equals the eager step, and a replay picks up
    expression values overwritten in place in the same device buffers."""
    import laris_b200 as la
    from laris_b200 import engine, _kernels as K
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(2, n=6000, G=200, G_u=120, P=70)
    xy_d = K.to_device(a.obsm["X_spatial"], torch.float64)
    csr = engine.upload_csr(a.X)
    lg, rg = la.tl._lr_gene_indices(a, lr)
    plan = engine.plan_lr_table(lg, rg, a.X.shape[1])

    def eager():
        g = engine.build_graph(xy_d, c.k, True, None)
        return engine.lr_scores(engine.select_planned(plan, csr, c.n, layout), g, plan.lig_d, plan.rec_d)

    ref = eager()
    cap = engine.CapturedScoring(xy_d, csr, plan, c.k, layout)
    assert cap.launches >= 10
    for _ in range(2):
        sm = cap.replay()
        torch.cuda.synchronize()
        assert torch.equal(sm.S, ref.S) and torch.equal(sm.row_nnz, ref.row_nnz)
    csr[2].mul_(2.0)                                     # same sparsity pattern, new values
    ref2 = eager()
    sm = cap.replay()
    torch.cuda.synchronize()
    assert torch.equal(sm.S, ref2.S) and torch.equal(ref2.S, ref.S * 4.0)


def test_compact_and_single_pass_selection_agree(dev, monkeypatch):
    """The sparse layout's two gene-selection forms (single pass with row ranges / compact two-pass CSR, used when the
    single-pass scratch would be too large) give bit-identical scores and COSG sums."""
    import laris_b200 as la
    from laris_b200 import engine, _kernels as K
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(3, n=4000, G=600, G_u=300, P=150)
    lg, rg = la.tl._lr_gene_indices(a, lr)
    g = engine.build_graph(a.obsm["X_spatial"], c.k, True, None)
    labels = K.to_device(a.obs["CellTypes"].cat.codes.to_numpy().astype(np.int32), torch.int32)
    res = []
    for cap in (engine.SINGLE_PASS_MAX_BYTES, 0):
        monkeypatch.setattr(engine, "SINGLE_PASS_MAX_BYTES", cap)
        xu, lc, rc = engine.select_lr_genes(a.X, lg, rg, layout="sparse")
        assert (xu.row_end is None) == (cap == 0)
        sm = engine.lr_scores(xu, g, lc, rc)
        s, cnt, q = K.onehot_contract_csr(xu.indptr, xu.packed, xu.G_u, labels, c.C, row_end=xu.row_end)
        res.append((sm.S.clone(), sm.row_nnz.clone(), cnt.clone(), q.clone(), s.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    assert torch.allclose(res[0][3], res[1][3], rtol=1e-12) and torch.allclose(res[0][4], res[1][4], rtol=1e-12)
