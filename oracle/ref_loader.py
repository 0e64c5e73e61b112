"""Load the unmodified reference ``laris.tools`` under stub modules.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works only where
``/root/reference`` is mounted (the build container); the GPU box does not
have it, so nothing on the ``-m gpu`` / smoke / bench paths may call this.

Recipe = SURVEY.md §10: ``scanpy``, ``anndata``, ``cosg``, ``statsmodels`` are
absent from the image, so tiny stand-ins are placed in ``sys.modules`` before
the reference's two numeric files are executed from where they lie.  The
package ``__init__`` (which imports the plotting stack) is never executed.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("LARIS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "laris", "tools", "_utils.py"))


def _install_stubs():
    from laris_b200._anndata import AnnDataLite

    def _bh(pvals, method="fdr_bh", alpha=0.05):
        # textbook Benjamini-Hochberg (statsmodels.stats.multitest.multipletests
        # 'fdr_bh'); statsmodels itself is absent from the image.
        p = np.asarray(pvals, dtype=np.float64)
        m = p.size
        order = np.argsort(p)
        ranked = p[order] * m / np.arange(1, m + 1)
        adj = np.minimum.accumulate(ranked[::-1])[::-1]
        adj = np.minimum(adj, 1.0)
        out = np.empty_like(adj)
        out[order] = adj
        return out <= alpha, out, None, None

    def _qc(a, inplace=True, percent_top=None):
        X = a.X.tocsr().copy()
        X.eliminate_zeros()
        a.var["n_cells_by_counts"] = np.asarray(X.getnnz(axis=0)).ravel()

    mods = {}
    if "anndata" not in sys.modules:
        m = types.ModuleType("anndata"); m.AnnData = AnnDataLite; mods["anndata"] = m
    if "scanpy" not in sys.modules:
        m = types.ModuleType("scanpy"); m.AnnData = AnnDataLite
        m.pp = types.SimpleNamespace(calculate_qc_metrics=_qc); mods["scanpy"] = m
    if "cosg" not in sys.modules:
        mods["cosg"] = _cosg_stub()
    if "statsmodels" not in sys.modules:
        sm = types.ModuleType("statsmodels"); st = types.ModuleType("statsmodels.stats")
        mt = types.ModuleType("statsmodels.stats.multitest"); mt.multipletests = _bh
        sm.stats = st; st.multitest = mt
        mods.update({"statsmodels": sm, "statsmodels.stats": st,
                     "statsmodels.stats.multitest": mt})
    sys.modules.update(mods)


def _cosg_stub():
    """Stand-in for the absent ``cosg`` package so that the reference's
    ``_calculate_laris_score_by_celltype`` can run end to end.  ``cosg.cosg`` is
    the CPU restatement (oracle.restate.cosg_scores, parity unpinned);
    ``indexByGene`` / ``iqrLogNormalize`` are the oracle's own stand-ins
    (oracle.cosg_standin, written independently of the product's
    laris_b200.cosg_compat), so everything AROUND them - table assembly,
    rescaling, the permutation null, BH-FDR - is pinned by the unmodified
    reference code, and the product's stand-ins are checked against them."""
    from . import cosg_standin, restate

    def cosg(adata, key_added="cosg", mu=1, expressed_pct=0.1, remove_lowly_expressed=False,
             n_genes_user=50, groupby="CellTypes", **_):
        lab = adata.obs[groupby]
        cats = list(lab.cat.categories)
        score = restate.cosg_scores(adata.X, lab.cat.codes.to_numpy(), len(cats), mu,
                                    expressed_pct, remove_lowly_expressed)
        frame = cosg_standin.ranked_frame(adata.var_names.to_numpy(), score, cats, "@@")
        adata.uns[key_added] = {"COSG": frame.iloc[:n_genes_user]}

    m = types.ModuleType("cosg")
    m.__laris_stub__ = True
    m.cosg = cosg
    m.indexByGene = lambda df, **kw: cosg_standin.index_by_gene(df, **kw)
    m.iqrLogNormalize = lambda df, **kw: cosg_standin.iqr_log_normalize(df, **kw)
    return m


_CACHE = {}


def load_reference_tools():
    """Return the reference's ``laris.tools`` module (with ``._utils``)."""
    if "tools" in _CACHE:
        return _CACHE["tools"]
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    pkg_dir = os.path.join(REFERENCE_ROOT, "laris")
    parent = types.ModuleType("laris_ref")
    parent.__path__ = [pkg_dir]
    sys.modules["laris_ref"] = parent
    tools_dir = os.path.join(pkg_dir, "tools")
    spec = importlib.util.spec_from_file_location(
        "laris_ref.tools", os.path.join(tools_dir, "__init__.py"),
        submodule_search_locations=[tools_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["laris_ref.tools"] = mod
    spec.loader.exec_module(mod)
    _CACHE["tools"] = mod
    return mod
