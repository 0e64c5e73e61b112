"""Minimal AnnData-compatible carrier.

The reference passes ``anndata.AnnData`` objects across its public boundary
(reference laris/tools/__init__.py:29-34, :121-153).  ``anndata`` is not
installed in this image, so the hot path accepts any object that quacks like
an AnnData (``X obs var obsm uns layers var_names obs_names n_obs n_vars shape
copy()``) and returns the real ``anndata.AnnData`` when it is importable and
this duck-typed carrier otherwise.  Only what the LR scoring path touches is
implemented; this is a data container, it does no arithmetic.
"""
from __future__ import annotations

import copy as _copy

import numpy as np
import pandas as pd
import scipy.sparse as sp


class AnnDataLite:
    """Duck-typed stand-in for ``anndata.AnnData`` (container only)."""

    def __init__(self, X=None, obs=None, var=None, obsm=None, uns=None, layers=None):
        if X is not None and not sp.issparse(X) and not getattr(X, "_laris_resident", False):
            X = np.asarray(X)
        self._X = X
        n_obs, n_vars = (X.shape if X is not None else (0, 0))
        if obs is None:
            obs = pd.DataFrame(index=pd.RangeIndex(n_obs).astype(str))
        if var is None:
            var = pd.DataFrame(index=pd.RangeIndex(n_vars).astype(str))
        if len(obs) != n_obs or len(var) != n_vars:
            raise ValueError(
                f"obs/var lengths ({len(obs)}, {len(var)}) do not match X shape {(n_obs, n_vars)}")
        self.obs = obs
        self.var = var
        self.obsm = dict(obsm) if obsm is not None else {}
        self.uns = dict(uns) if uns is not None else {}
        self.layers = dict(layers) if layers is not None else {}

    # -- matrix ---------------------------------------------------------
    @property
    def X(self):
        return self._X

    @X.setter
    def X(self, value):
        if value.shape != self.shape:
            raise ValueError(f"X shape {value.shape} does not match {self.shape}")
        self._X = value

    # -- shape ----------------------------------------------------------
    @property
    def shape(self):
        return (len(self.obs), len(self.var))

    @property
    def n_obs(self):
        return len(self.obs)

    @property
    def n_vars(self):
        return len(self.var)

    # -- names ----------------------------------------------------------
    @property
    def var_names(self):
        return self.var.index

    @var_names.setter
    def var_names(self, names):
        self.var.index = pd.Index(np.asarray(names, dtype=object), dtype=object)

    @property
    def obs_names(self):
        return self.obs.index

    @obs_names.setter
    def obs_names(self, names):
        self.obs.index = pd.Index(np.asarray(names, dtype=object), dtype=object)

    # -- copy / slice ---------------------------------------------------
    def copy(self):
        X = self._X.copy() if self._X is not None and hasattr(self._X, "copy") else self._X
        return type(self)(
            X, self.obs.copy(), self.var.copy(),
            {k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.obsm.items()},
            _copy.deepcopy(self.uns),
            {k: v.copy() for k, v in self.layers.items()},
        )

    def _axis_index(self, key, names, size):
        if isinstance(key, slice):
            return np.arange(size)[key]
        key = np.asarray(key) if not isinstance(key, (pd.Series, pd.Index)) else np.asarray(key)
        if key.dtype == bool:
            return np.flatnonzero(key)
        if key.dtype.kind in "iu":
            return key.astype(np.int64)
        idx = names.get_indexer(key)
        if (idx < 0).any():
            raise KeyError(f"names not found: {list(np.asarray(key)[idx < 0][:5])}")
        return idx

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key, slice(None))
        r = self._axis_index(key[0], self.obs_names, self.n_obs)
        c = self._axis_index(key[1], self.var_names, self.n_vars)
        X = self._X
        if X is not None:
            X = X.tocsr()[r][:, c] if sp.issparse(X) else X[np.ix_(r, c)]
        return type(self)(
            X, self.obs.iloc[r].copy(), self.var.iloc[c].copy(),
            {k: np.asarray(v)[r] for k, v in self.obsm.items()},
            _copy.deepcopy(self.uns),
            {k: (v.tocsr()[r][:, c] if sp.issparse(v) else v[np.ix_(r, c)])
             for k, v in self.layers.items()},
        )

    def __repr__(self):
        return (f"AnnDataLite n_obs x n_vars = {self.n_obs} x {self.n_vars}; "
                f"obs={list(self.obs.columns)} var={list(self.var.columns)} "
                f"obsm={list(self.obsm)} uns={list(self.uns)}")


def make_anndata(X, obs=None, var=None, obsm=None, uns=None):
    """Return a real ``anndata.AnnData`` when importable, else :class:`AnnDataLite`."""
    try:
        import anndata  # noqa: F401
    except ImportError:
        return AnnDataLite(X, obs, var, obsm, uns)
    a = anndata.AnnData(X)
    if obs is not None:
        a.obs = obs
    if var is not None:
        a.var = var
    for k, v in (obsm or {}).items():
        a.obsm[k] = v
    for k, v in (uns or {}).items():
        a.uns[k] = v
    return a
