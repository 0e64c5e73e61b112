"""Property tests (hypothesis) of the CUDA path against the oracle on tiny random inputs, plus size-independent
invariants of the scoring path (SURVEY.md §7.3 test pyramid, tier iii).  Every example goes through the C ABI."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, HealthCheck, strategies as st

pytestmark = pytest.mark.gpu
import os
COMMON = dict(deadline=None, max_examples=int(os.environ.get("LARIS_HYP_EXAMPLES", "30")), suppress_health_check=list(HealthCheck), derandomize=not os.environ.get("LARIS_HYP_RANDOM"), database=None)
RTOL = 1e-5


def _points(rng, n, dim, kind):
    if kind == 0:
        return rng.random((n, dim)) * 100.0
    if kind == 1:                                   # clustered, very uneven density
        return rng.normal(0, 1, (n, dim)) * rng.choice([0.01, 1.0, 30.0], (n, 1))
    xy = rng.random((n, dim)) * 100.0               # anisotropic extent
    xy[:, 1] *= 1e-3
    return xy


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(8, 200), k=st.integers(1, 15), dim=st.sampled_from([2, 3]),
       include_self=st.booleans(), kind=st.integers(0, 2))
def test_knn_matches_sklearn_on_random_points(seed, n, k, dim, include_self, kind):
    from laris_b200 import engine
    from oracle import restate as R
    k = min(k, n // 2 - 2)                          # sklearn goes brute force (GEMM distances, not bit-reproducible) at k >= n // 2
    rng = np.random.default_rng(seed)
    xy = _points(rng, n, dim, kind)
    g = engine.build_graph(xy, k, include_self, sigma=50.0)
    ind, dist = R.knn_graph(xy, k, include_self)
    assert np.array_equal(g.dist.cpu().numpy(), dist)
    assert np.array_equal(g.idx.cpu().numpy(), ind)
    assert np.array_equal(np.sort(g.order.cpu().numpy()), np.arange(n))


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 200), dim=st.sampled_from([2, 3]), include_self=st.booleans(),
       q=st.floats(0.02, 0.6))
def test_radius_graph_matches_sklearn_on_random_points(seed, n, dim, include_self, q):
    from laris_b200 import _kernels as K
    from oracle import restate as R
    rng = np.random.default_rng(seed)
    xy = _points(rng, n, dim, seed % 3)
    radius = float(q * (np.ptp(xy, axis=0).max() + 1e-3))
    indptr, indices, dist, _ = K.radius_graph(K.to_device(xy, torch.float64), radius, include_self)
    r_indptr, r_indices, r_dist = R.radius_graph(xy, radius, include_self)
    assert np.array_equal(indptr.cpu().numpy(), r_indptr)
    assert np.array_equal(indices.cpu().numpy(), r_indices)
    if n >= 12:
        assert np.array_equal(dist.cpu().numpy(), r_dist)
    else:       # NearestNeighbors' default n_neighbors=5 >= n//2 sends sklearn to brute force (GEMM-expanded distances)
        assert np.allclose(dist.cpu().numpy(), r_dist, rtol=1e-9, atol=1e-9)


def _tiny_problem(rng, n, G, P, density, signed=False):
    from laris_b200._anndata import make_anndata
    X = sp.random(n, G, density=density, format="csr", random_state=np.random.RandomState(int(rng.integers(1 << 30))),
                  data_rvs=lambda m: np.log1p(1.0 + rng.poisson(2.0, m)).astype(np.float32)).astype(np.float32)
    if signed:
        X.data[::3] *= -1.0
    names = np.array([f"g{j:03d}" for j in range(G)], dtype=object)
    lr = pd.DataFrame({"ligand": names[rng.integers(0, G, P)], "receptor": names[rng.integers(0, G, P)]})
    a = make_anndata(X, obs=pd.DataFrame(index=[f"c{i}" for i in range(n)]), var=pd.DataFrame(index=names),
                     obsm={"X_spatial": rng.random((n, 2)) * 50.0})
    return a, lr


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(24, 200), G=st.integers(2, 40), P=st.integers(1, 30),
       k=st.integers(2, 10), density=st.floats(0.02, 0.9), layout=st.sampled_from(["dense", "sparse"]), signed=st.booleans())
def test_scores_match_oracle_on_random_problems(seed, n, G, P, k, density, layout, signed):
    import laris_b200 as la
    from oracle import restate as R
    rng = np.random.default_rng(seed)
    a, lr = _tiny_problem(rng, n, G, P, density, signed)
    out = la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=k, layout=layout)
    ind, dist = R.knn_graph(a.obsm["X_spatial"], k, True)
    S = R.lr_scores(a.X, ind, R.decay_weights(dist), R.gene_index(a.var_names, lr["ligand"]),
                    R.gene_index(a.var_names, lr["receptor"]))
    X = sp.csr_matrix(out.X)
    X.sort_indices()
    if not signed:                                  # non-negative data: structure is exact
        assert np.array_equal(X.indptr, S.indptr) and np.array_equal(X.indices, S.indices)
        assert np.allclose(X.data, S.data, rtol=RTOL, atol=0)
    else:                                           # mixed signs: a sum may cancel to (almost) zero
        assert np.allclose(X.toarray(), S.toarray(), rtol=1e-4, atol=1e-5)


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(24, 200), P=st.integers(1, 40), k=st.integers(1, 12),
       l1=st.booleans())
def test_spatial_statistics_match_oracle(seed, n, P, k, l1):
    from laris_b200 import _kernels as K
    from oracle import restate as R
    rng = np.random.default_rng(seed)
    S = ((rng.random((n, P)) < 0.4) * rng.gamma(2.0, 1.0, (n, P))).astype(np.float32)
    idx = rng.integers(0, n, (n, k)).astype(np.int32)           # arbitrary graph, duplicates allowed
    w = rng.random((n, k)).astype(np.float32)
    Sd = torch.zeros((n, K.padded_ld(P)), dtype=torch.float32, device="cuda")
    Sd[:, :P] = torch.from_numpy(S).cuda()
    st5 = K.spatial_stats(Sd, P, torch.from_numpy(idx).cuda(), torch.from_numpy(w).cuda(), l1, None).cpu().numpy()
    W = sp.csr_matrix((w.ravel().astype(np.float64), (np.repeat(np.arange(n), k), idx.ravel())), shape=(n, n))
    if l1:
        from sklearn.preprocessing import normalize
        W = normalize(W, axis=1, norm="l1")
    S64 = S.astype(np.float64)
    T = W @ S64
    assert np.allclose(st5[0], (S64 * T).sum(0), rtol=2e-5, atol=1e-9)
    assert np.allclose(st5[1], (S64 * S64).sum(0), rtol=1e-6)
    assert np.allclose(st5[2], (T * T).sum(0), rtol=2e-5, atol=1e-9)
    assert np.allclose(st5[3], S64.sum(0), rtol=1e-6)
    assert np.array_equal(st5[4], (S != 0).sum(0))


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31 - 1), groups=st.integers(1, 12), P=st.integers(31, 80), N=st.integers(1, 300))
def test_pvalue_counts_and_fdr_match_oracle(seed, groups, P, N):
    from laris_b200 import engine
    from oracle import restate as R, rng as orng
    rng = np.random.default_rng(seed)
    table = rng.random((groups, P)) * (rng.random((groups, P)) < 0.7)
    active = (rng.random(P) < 0.8).astype(np.uint8)
    table[:, active == 0] = 0.0                     # pairs outside the ranked list carry 0.0 (reference _utils.py:1563)
    bg = np.stack([rng.permutation(P)[:30] for _ in range(P)]).astype(np.int32)
    draws = rng.integers(0, 30, (groups, P, N)).astype(np.uint8)
    ge, nz, both = engine.pvalue_counts(table, bg, active, N, rng="numpy", draws=draws)
    m = active.astype(bool)
    rge, rnz, rboth = R.pvalue_counts(table, m, bg.astype(np.int64), lambda g, pidx: draws[g][pidx].astype(np.int64))
    assert np.array_equal(ge[:, m], rge[:, m]) and np.array_equal(nz[:, m], rnz[:, m]) and np.array_equal(both[:, m], rboth[:, m])
    p = np.round((ge + 1) / (N + 1), 6)
    sel = np.broadcast_to(m[None, :], table.shape) & (table > 0)
    fdr = engine.bh_fdr(p, sel.astype(np.uint8))
    for g in range(groups):
        if sel[g].any():
            assert np.allclose(fdr[g][sel[g]], R.bh_fdr(p[g][sel[g]]), rtol=1e-12)
        assert (fdr[g][~sel[g]] == 1.0).all()


# ------------------------------------------------ size-independent invariants ----
def test_scores_are_quadratic_in_expression_and_relabelling_invariant():
    """S(aX) = a^2 S(X) for a power of two (exact), and renumbering the cells only permutes the rows of S."""
    import laris_b200 as la
    from laris_b200.synth import make_dataset
    a, lr, c = make_dataset(2, n=30000)
    base = sp.csr_matrix(la.tl.prepareLRInteraction(a, lr, number_nearest_neighbors=c.k).X)
    b = a.copy()
    b.X = (a.X * np.float32(4.0)).tocsr()
    quad = sp.csr_matrix(la.tl.prepareLRInteraction(b, lr, number_nearest_neighbors=c.k).X)
    assert np.array_equal(quad.indptr, base.indptr) and np.array_equal(quad.indices, base.indices)
    assert np.array_equal(quad.data, base.data * np.float32(16.0))           # powers of two scale exactly
    perm = np.random.default_rng(1).permutation(c.n)
    p = a.copy()
    p.X = a.X[perm].tocsr()
    p.obs = a.obs.iloc[perm]
    p.obsm["X_spatial"] = a.obsm["X_spatial"][perm]
    shuf = sp.csr_matrix(la.tl.prepareLRInteraction(p, lr, number_nearest_neighbors=c.k).X)
    ref = base[perm]
    ref.sort_indices()
    shuf.sort_indices()
    assert np.array_equal(shuf.indptr, ref.indptr) and np.array_equal(shuf.indices, ref.indices)
    assert np.allclose(shuf.data, ref.data, rtol=2e-6, atol=0)              # the dense kernel sums a cell's rows in warp-group order


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 3000), P=st.integers(1, 300), C=st.integers(1, 40),
       dens=st.floats(0.01, 1.0), types=st.integers(1, 4), perm=st.booleans())
def test_tensor_core_contraction_on_random_shapes(seed, n, P, C, dens, types, perm):
    """tcgen05 pair contraction (three-way bf16 split) against a float64 contraction: 1e-5 on every non-zero entry,
    exact zeros stay exact - ragged tiles, single cells, dense and very sparse score matrices, any cell order."""
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(seed)
    ld = K.padded_ld(P)
    S = np.zeros((n, ld), np.float32)
    S[:, :P] = (rng.random((n, P)) < dens) * rng.gamma(2.0, 1.0, (n, P)).astype(np.float32) * \
        np.float32(10.0) ** rng.integers(-3, 4, (1, P)).astype(np.float32)          # columns of very different scale
    N = np.zeros((n, C), np.float32)
    for _ in range(types):
        N[np.arange(n), rng.integers(0, C, n)] += rng.random(n).astype(np.float32)
    Sd, Nd = torch.from_numpy(S).cuda(), torch.from_numpy(N).cuda()
    order = torch.from_numpy(rng.permutation(n).astype(np.int32)).cuda() if perm else None
    num, q = K.celltype_pair_contract(Sd, P, Nd, 0, order=order)
    Q = (N[:, :, None].astype(np.float64) * N[:, None, :]).reshape(n, C * C)
    ref = S[:, :P].astype(np.float64).T @ Q
    got = num.cpu().numpy()
    assert np.array_equal(got == 0, ref == 0)
    assert np.allclose(got, ref, rtol=RTOL, atol=0)
    assert np.allclose(q.cpu().numpy(), (Q * Q).sum(0), rtol=1e-6)


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 2000), P=st.integers(1, 200), C=st.integers(1, 60))
def test_celltype_counts_and_cell_qc_on_random_shapes(seed, n, P, C):
    from laris_b200 import _kernels as K
    rng = np.random.default_rng(seed)
    S = np.zeros((n, K.padded_ld(P)), np.float32)
    S[:, :P] = (rng.random((n, P)) < 0.3) * rng.random((n, P)).astype(np.float32)
    labels = rng.integers(0, C, n).astype(np.int32)
    Sd = torch.from_numpy(S).cuda()
    got = K.celltype_nonzero_count(Sd, P, torch.from_numpy(labels).cuda(), C).cpu().numpy()
    ref = np.zeros((C, P), np.int64)
    np.add.at(ref, labels, (S[:, :P] != 0).astype(np.int64))
    assert np.array_equal(got, ref)
    tops = [t for t in (1, 7, 50) if t < P] or [1]
    total, nnz, top = K.row_qc(Sd, P, tops)
    r = S[:, :P].astype(np.float64)
    srt = -np.sort(-r, axis=1)
    assert np.allclose(total.cpu().numpy(), r.sum(1), rtol=1e-12, atol=1e-12)
    assert np.array_equal(nnz.cpu().numpy(), (r != 0).sum(1))
    for j, t in enumerate(tops):
        assert np.allclose(top.cpu().numpy()[:, j], srt[:, :min(t, P)].sum(1), rtol=1e-12, atol=1e-12)
