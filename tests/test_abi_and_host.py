"""CPU tier: the C-ABI library loads and exports every symbol the header
declares; host-side logic (carrier, generators, table reshaping, argument
errors) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "laris_b200.h")).read()
    declared = set(re.findall(r"LARIS_API[\s\w\*]+?\b(laris_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "laris_b200", "liblaris_b200.so"))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from laris_b200 import _lib
    assert declared - {"laris_last_error", "laris_launch_count"} == set(_lib.SIGNATURES), "ctypes table out of sync with the header"
    assert _lib.lib.laris_abi_version() == 1


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "laris_b200", "liblaris_b200.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cuda_means_loud_failure(tiny):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import laris_b200 as la
    a, lr, _ = tiny
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        la.tl.prepareLRInteraction(a, lr)


def test_product_never_imports_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "laris_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f


def test_argument_errors_match_reference(tiny):
    import laris_b200 as la
    a, lr, _ = tiny
    lr_adata = a[:, :5]
    with pytest.raises(ValueError, match="must be provided when by_celltype=True"):
        la.tl.runLARIS(lr_adata, None, by_celltype=True)                       # reference tools/__init__.py:424-428
    with pytest.raises(KeyError, match="not found in lr_adata.obsm"):
        la.tl.runLARIS(lr_adata, a, use_rep="nope")                            # :430-434
    with pytest.raises(ValueError, match="same shape"):
        la.pp._rowwise_cosine_similarity(np.zeros((2, 3)), np.zeros((3, 2)))   # _utils.py:81-82
    with pytest.raises(ValueError, match="must match number of rows"):
        la.pp._pairwise_row_multiply(sp.csr_matrix(np.eye(3)), ["a", "b"])     # _utils.py:211-212
    with pytest.raises(ValueError):
        la.pp._select_top_n(np.arange(5.0), 9)                                 # argpartition kth out of bounds
    bad = pd.DataFrame({"ligand": ["g00001", "nope"], "receptor": ["g00002", "g00003"]})
    with pytest.raises(KeyError, match="nope"):
        la.tl._lr_gene_indices(a, bad)                                         # FIX of the silent mis-map (:95-97)


def test_select_top_n_docstring():
    import laris_b200 as la
    assert list(la.pp._select_top_n(np.array([0.1, 0.9, 0.3, 0.7, 0.5]), 3)) == [1, 3, 4]


def test_anndata_carrier(tiny):
    a, _, c = tiny
    assert a.shape == (c.n, c.G) and a.n_obs == c.n and a.n_vars == c.G
    b = a[:, ["g00003", "g00001"]]
    assert b.shape == (c.n, 2) and list(b.var_names) == ["g00003", "g00001"]
    assert (b.X != sp.csr_matrix(a.X)[:, [3, 1]]).nnz == 0
    m = (a.obs["CellTypes"] == "ct01").to_numpy()
    assert a[m, :].n_obs == m.sum()
    d = a.copy()
    d.obsm["X_spatial"][0, 0] = -1
    assert a.obsm["X_spatial"][0, 0] != -1


def test_synthetic_configs():
    from laris_b200.synth import CONFIGS, make_dataset
    assert set(CONFIGS) == {1, 2, 3, 4, 5}
    a, lr, c = make_dataset(3, n=2000, G=200, G_u=120, P=90)
    assert lr.shape == (90, 2) and not lr.duplicated().any()
    assert lr["ligand"].value_counts().max() >= 2          # complexes expanded to several rows per ligand
    xy = a.obsm["X_spatial"]
    assert len(np.unique(xy, axis=0)) == len(xy)
    b, _, _ = make_dataset(3, n=2000, G=200, G_u=120, P=90)
    assert np.array_equal(b.obsm["X_spatial"], xy) and (b.X != a.X).nnz == 0


def test_ranked_frame_roundtrip():
    from laris_b200 import cosg_compat
    rng = np.random.default_rng(0)
    sc = rng.random((7, 3))
    sc[2, 1] = -1.0
    names = np.array([f"n{i}" for i in range(7)], dtype=object)
    frame, order = cosg_compat.ranked_frame(names, sc, ["a", "b", "c"], "@@")
    assert list(frame.columns[:3]) == ["names@@a", "names@@b", "names@@c"]
    assert (np.diff(frame["scores@@a"].to_numpy()) <= 0).all()
    back = cosg_compat.index_by_gene(frame, set_nan_to_zero=True, convert_negative_one_to_zero=True).loc[names]
    expect = sc.copy()
    expect[2, 1] = 0.0
    assert np.allclose(back.to_numpy(), expect)
    norm = cosg_compat.iqr_log_normalize(back)
    assert norm.shape == back.shape and (norm.to_numpy() >= 0).all() and norm.to_numpy()[2, 1] == 0


def test_lr_column_layout_is_injective_and_conflict_free():
    """Host helper laris_lr_column_layout: an injective column numbering under which the score
    kernel's 32-lane gathers (pairs 128b + 4t + u, t = lane) hit 32 distinct banks."""
    from laris_b200 import engine
    rng = np.random.default_rng(3)
    for G_u, P in ((40, 64), (334, 500), (1346, 2000), (7, 300)):
        lig = rng.integers(0, G_u, P).astype(np.int32)
        rec = rng.integers(0, G_u, P).astype(np.int32)
        cols, ncol = engine.lr_column_layout(lig, rec, G_u)
        assert cols.shape == (G_u,) and len(np.unique(cols)) == G_u and cols.min() >= 0 and cols.max() < ncol
        assert ncol % 32 == 0 and ncol <= G_u + 96
        got = ident = groups = 0
        for a in (lig, rec):
            for b0 in range(0, P, 128):
                for u in range(4):
                    g = np.unique(a[b0 + u:b0 + 128:4])
                    got += np.bincount(cols[g] % 32).max()          # wavefronts of that gather
                    ident += np.bincount(g % 32).max()
                    groups += 1
        assert got <= ident and (G_u < 64 or got <= 0.7 * ident), (got, ident, groups)


def test_complex_name_parsing(tiny):
    """Host side of the multi-subunit extension: names with the separator become complexes (ids >= n_genes), a
    trailing separator is just the gene, unknown subunits raise."""
    import laris_b200 as la
    a, _, _ = tiny
    G = a.shape[1]
    names = list(a.var_names)
    lr = pd.DataFrame({"ligand": [names[3], f"{names[4]}+{names[5]}", names[3], f"{names[6]}+"],
                       "receptor": [f"{names[7]}+{names[8]}+{names[9]}", names[10], f"{names[4]}+{names[5]}", names[11]]})
    lig, rec, complexes = la.tl._lr_units(a, lr, "+")
    assert lig[0] == 3 and lig[2] == 3 and rec[1] == 10 and rec[3] == 11 and lig[3] == 6       # "g+" is the gene itself
    assert lig[1] >= G and rec[0] >= G and lig[1] == rec[2]                                     # one id per distinct complex
    assert sorted(len(c) for c in complexes) == [2, 3]
    assert list(complexes[lig[1] - G]) == [4, 5] and list(complexes[rec[0] - G]) == [7, 8, 9]
    bad = lr.copy()
    bad.loc[0, "receptor"] = f"{names[7]}+missing_gene"
    with pytest.raises(KeyError, match="missing_gene"):
        la.tl._lr_units(a, bad, "+")


def test_index_by_gene_is_a_pivot():
    """cosg_compat.index_by_gene (vectorised) equals the per-column definition on a ranked frame with ties and -1."""
    from laris_b200 import cosg_compat
    rng = np.random.default_rng(3)
    names = np.array([f"n{i}" for i in range(40)], dtype=object)
    scores = rng.random((40, 6))
    scores[rng.random((40, 6)) < 0.2] = -1.0
    groups = [f"a::{j}" for j in range(6)]
    frame, order = cosg_compat.ranked_frame(names, scores, groups, "@@")
    assert all((np.diff(frame[f"scores@@{g}"].to_numpy()) <= 0).all() for g in groups)
    table = cosg_compat.index_by_gene(frame, column_delimiter="@@")
    want = pd.DataFrame(np.where(scores == -1, 0.0, scores), index=names, columns=groups)
    assert np.allclose(table.loc[names, groups].to_numpy(), want.to_numpy())


def test_shard_bounds_and_layout_choice():
    from laris_b200 import dist as ld, engine
    assert list(ld.pair_bounds(10, 4)) == [0, 3, 6, 8, 10]
    X = sp.random(100, 50, density=0.3, format="csr", random_state=1)
    assert engine.choose_layout(X) == "dense" and engine.choose_layout(X, "sparse") == "sparse"
    assert engine.choose_layout(sp.random(100, 50, density=0.05, format="csr", random_state=1)) == "sparse"


def test_take_strings_equals_object_gather():
    from laris_b200.tools._utils import _take_strings
    vals = np.array(["ct00", "ct01", "b::c", "é"], dtype=object)
    codes = np.random.default_rng(0).integers(0, 4, 1000)
    a = pd.DataFrame({"x": _take_strings(vals, codes)})["x"]
    b = pd.DataFrame({"x": vals[codes]})["x"]
    assert a.dtype == b.dtype and a.equals(b)
    cats = np.array([3, 1, 2], dtype=object)                       # non-string categories: plain gather
    assert list(_take_strings(cats, np.array([2, 0]))) == [2, 3]
    assert len(_take_strings(np.array([], dtype=object), np.array([], dtype=np.int64))) == 0


def test_pivot_of_a_ranking_equals_index_by_gene():
    from laris_b200 import cosg_compat
    rng = np.random.default_rng(5)
    names = np.array([f"n{i}" for i in range(60)], dtype=object)
    scores = rng.random((60, 7))
    scores[rng.random((60, 7)) < 0.15] = -1.0
    groups = [f"g{j}" for j in range(7)]
    frame, order = cosg_compat.ranked_frame(names, scores, groups, "@@")
    for kw in (dict(), dict(set_nan_to_zero=True, convert_negative_one_to_zero=True), dict(convert_negative_one_to_zero=False)):
        a = cosg_compat.index_by_gene(frame, **kw)
        b = cosg_compat.index_by_gene_of_ranking(frame, names, scores, groups, order, **kw)
        assert list(a.index) == list(b.index) and list(a.columns) == list(b.columns)
        assert np.array_equal(a.to_numpy(), b.to_numpy())


def test_signatures_match_the_reference():
    """Drop-in boundary: every function the reference defines in its hot-path modules exists here with the same
    positional parameters (names, order, defaults); anything extra is keyword-only.  The list was generated from the
    unmodified reference by oracle/make_signatures.py."""
    import ast
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_signatures.json")))
    for rel, spec in ref.items():
        tree = ast.parse(open(os.path.join(ROOT, "laris_b200", rel)).read())
        ours = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
        for name, sig in spec["functions"].items():
            assert name in ours, f"{rel}: reference function {name} is missing"
            a = ours[name].args
            pos = [x.arg for x in a.posonlyargs + a.args]
            defaults = [ast.unparse(d) for d in a.defaults]
            dmap = dict(zip(pos[len(pos) - len(defaults):], defaults))
            got = [[p, dmap.get(p)] for p in pos]
            want = [[p, d] for p, d in sig["positional"]]
            norm = lambda s: None if s is None else s.replace('"', "'")        # noqa: E731
            assert [[p, norm(d)] for p, d in got] == [[p, norm(d)] for p, d in want], f"{rel}:{name} signature drifted"
    import laris_b200 as la
    for name in ref["tools/__init__.py"]["__all__"]:
        assert hasattr(la.tl, name)
    for name in ref["tools/_utils.py"]["__all__"]:
        assert hasattr(la.tl._utils, name)
    for name in ref["preprocessing/__init__.py"]["__all__"]:
        assert hasattr(la.pp, name)


def test_product_stand_ins_equal_the_oracle_stand_ins():
    """laris_b200.cosg_compat (product) and oracle.cosg_standin (test infrastructure) are written independently;
    on random tables with ties, zeros, -1 marks and empty columns they must agree."""
    from laris_b200 import cosg_compat
    from oracle import cosg_standin
    rng = np.random.default_rng(11)
    names = np.array([f"n{i}" for i in range(57)], dtype=object)
    scores = rng.random((57, 6)) ** 3
    scores[rng.random((57, 6)) < 0.3] = 0.0
    scores[rng.random((57, 6)) < 0.1] = -1.0
    scores[:, 4] = 0.0                                  # a column without positive entries
    scores[:5, 5] = scores[5, 5]                        # ties
    groups = [f"g{j}" for j in range(6)]
    f1, _ = cosg_compat.ranked_frame(names, scores, groups, "@@")
    f2 = cosg_standin.ranked_frame(names, scores, groups, "@@")
    assert list(f1.columns) == list(f2.columns)
    for c in f1.columns:
        assert list(f1[c]) == list(f2[c]), c
    t1 = cosg_compat.index_by_gene(f1, set_nan_to_zero=True)
    t2 = cosg_standin.index_by_gene(f2, set_nan_to_zero=True)
    assert list(t1.index) == list(t2.index) and np.array_equal(t1.to_numpy(), t2.to_numpy())
    n1 = cosg_compat.iqr_log_normalize(t1).to_numpy()
    n2 = cosg_standin.iqr_log_normalize(t2).to_numpy()
    assert np.allclose(n1, n2, rtol=1e-14, atol=0) and n1.max() == 1.0 and n1.min() == 0.0 and (n1[:, 4] == 0).all()


def test_pending_csr_is_a_scipy_matrix_that_joins_once():
    """engine.PendingCSR without a GPU: the arrays are joined on first access, exactly once; copies / pickles are plain."""
    import pickle
    from laris_b200 import engine
    A = sp.random(40, 17, 0.3, format="csr", dtype=np.float32, random_state=3)
    X = engine.PendingCSR((40, 17), dtype=np.float32)
    X.indptr, X.indices, X.data = A.indptr.copy(), A.indices.copy(), A.data.copy()
    X.has_sorted_indices = X.has_canonical_format = True
    joined = []
    X.__dict__["_laris_wait"] = lambda m: joined.append(m.shape)
    assert X._laris_pending() and X.shape == (40, 17) and sp.issparse(X) and not joined      # shape alone does not join
    assert X.nnz == A.nnz and joined == [(40, 17)] and not X._laris_pending()
    assert (X != A).nnz == 0 and np.allclose((X @ np.ones(17)), A @ np.ones(17)) and len(joined) == 1
    assert type(pickle.loads(pickle.dumps(X))) is sp.csr_matrix and (X.copy() != A).nnz == 0
    X.data *= 2.0
    assert np.allclose(X.sum(), 2.0 * A.sum())


def test_take_many_equals_single_takes():
    from laris_b200 import cosg_compat
    rng = np.random.default_rng(1)
    names = np.array([f"n{i}" for i in range(50)], dtype=object)
    codes = [rng.integers(0, 50, 300_000).astype(np.int32), rng.integers(0, 50, 300_000)]
    t = cosg_compat.StringTaker(names)
    many = cosg_compat.take_many([(t, codes[0]), (t, codes[1])])
    for got, c in zip(many, codes):
        assert (np.asarray(got, dtype=object) == names[c]).all()


def test_table_row_order_mirror_equals_the_oracle_sorts():
    """``_rows_in_table_order`` (the host mirror and test oracle of the device ordering, T6a) against the oracle's own
    restatement of reference _utils.py:1408-1422, :1472: melt order, stable sort by the pre-weight score, stable sort by
    the final score.  The rows that keep a positive score come out in exactly the oracle's order (ties included); the
    thresholded-away rows follow in melt order."""
    from laris_b200.tools import _utils as U
    rng = np.random.default_rng(11)
    for npair, C in ((5, 3), (40, 6), (3, 1)):
        L = C * C
        inter = np.round(rng.random((npair, C, C)), 1)                        # [pair, sender, receiver], many ties
        weight = np.round(rng.random((npair, C, C)), 1)
        final = inter * weight
        final[final < 0.2] = 0.0                                              # the threshold of :1480-1489
        # oracle/restate.py run_step2: melt [pair, receiver, sender], o1 by -inter, o2 by -(inter * weight)
        melt_inter = inter.transpose(0, 2, 1).reshape(-1)
        elem = np.arange(npair * L).reshape(npair, C, C).transpose(0, 2, 1).reshape(-1)
        o1 = np.argsort(-melt_inter, kind="stable")
        prod = (inter * weight).reshape(-1)[elem[o1]]
        o2 = np.argsort(-prod, kind="stable")
        oracle_rows = elem[o1][o2]
        oracle_rows = oracle_rows[final.reshape(-1)[oracle_rows] > 0]
        got = U._rows_in_table_order(inter.reshape(npair, L), final.reshape(npair, L), npair, C)
        npos = int((final > 0).sum())
        assert sorted(got.tolist()) == list(range(npair * L))
        assert np.array_equal(got[:npos], oracle_rows)
        rest = got[npos:]
        assert (final.reshape(-1)[rest] == 0).all()
        assert np.array_equal(rest, elem[final.reshape(-1)[elem] == 0])


def test_compute_avg_expression_matches_the_reference(golden, tiny):
    """Host helper laris/tools/_utils.py:248-340 against outputs of the unmodified reference (oracle/make_golden.py): per-category
    column means, categories x genes, gene subsets in var order; same ValueErrors.  (The reference averages float32 sparse rows
    in float32; here the sums are float64 - hence the 1e-5 tolerance.)"""
    import warnings
    from laris_b200.tools import _utils as U
    a, _, _ = tiny
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        full = U._compute_avg_expression(a, groupby="CellTypes")
        sub = U._compute_avg_expression(a, groupby="CellTypes", genes=list(golden["avg_gene_query"]))
        one = U._compute_avg_expression(a, groupby="CellTypes", groups=["ct03", "ct01"])
    assert list(full.index) == list(a.obs["CellTypes"].cat.categories) and list(full.columns) == list(a.var_names)
    assert np.allclose(full.to_numpy(dtype=np.float64), golden["avg_all"], rtol=1e-5, atol=1e-9)
    assert list(sub.columns) == list(golden["avg_gene_columns"])
    assert np.allclose(sub.to_numpy(dtype=np.float64), golden["avg_genes"], rtol=1e-5, atol=1e-9)
    assert list(one.index) == ["ct01", "ct03"] and np.allclose(one.to_numpy(dtype=np.float64), golden["avg_all"][[1, 3]], rtol=1e-5)
    for bad in (dict(groupby="nope"), dict(groupby="CellTypes", genes=["zzz"]), dict(groupby="CellTypes", groups=["zzz"])):
        with pytest.raises(ValueError):
            U._compute_avg_expression(a, **bad)


def test_pending_csr_concurrent_readers_all_wait_for_the_join():
    """Two threads touching a pending matrix at once: the second waits until the first has joined the copies."""
    import threading
    import time
    from laris_b200 import engine
    A = sp.random(30, 11, 0.4, format="csr", dtype=np.float32, random_state=5)
    X = engine.PendingCSR((30, 11), dtype=np.float32)
    X.indptr, X.indices, X.data = A.indptr.copy(), A.indices.copy(), np.zeros_like(A.data)      # "not arrived yet"
    joins = []

    def wait(m):
        time.sleep(0.2)
        m.__dict__["_laris_data"][:] = A.data
        joins.append(1)
    X.__dict__["_laris_wait"] = wait
    seen = []
    threads = [threading.Thread(target=lambda: seen.append(float(X.data.sum()))) for _ in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert joins == [1] and seen == [float(A.data.sum())] * 4 and not X._laris_pending()
