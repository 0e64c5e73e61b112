"""Seeded synthetic spatial datasets for the five BASELINE.json configs.

Shapes, densities and the label model follow SURVEY.md §8(d).  Everything is
generated with ``np.random.default_rng(1000 + cfg)`` so tests, the oracle and
``bench.py`` see the same bytes.  All coordinate sets are tie-free (jittered),
which is the only regime where the reference's kNN graph is itself
well-defined.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd
import scipy.sparse as sp

from ._anndata import make_anndata


@dataclass(frozen=True)
class SynthConfig:
    name: str
    n: int          # cells / spots / bins
    G: int          # genes
    G_u: int        # distinct genes drawn on by the LR table
    P: int          # LR pairs (rows of lr_df, complexes pre-expanded)
    k: int          # neighbours
    C: int          # cell types
    R: int          # random-graph repeats
    N: int          # p-value permutations
    density: float  # mean fraction of non-zeros in X
    coords: str     # 'hex' | 'uniform' | 'bins'
    complex_frac: float = 0.0


CONFIGS = {
    1: SynthConfig("visium_like", 10_000, 2_000, 600, 500, 10, 10, 3, 100, 0.10, "hex"),
    2: SynthConfig("merfish_like", 100_000, 500, 350, 500, 10, 20, 3, 1000, 0.20, "uniform"),
    3: SynthConfig("xenium_like", 1_000_000, 5_000, 1_500, 2_000, 10, 40, 3, 1000, 0.10, "uniform", 0.2),
    4: SynthConfig("cosmx_like", 5_000_000, 1_000, 600, 1_000, 10, 40, 3, 1000, 0.10, "uniform"),
    5: SynthConfig("stereoseq_bins", 10_000_000, 2_000, 500, 500, 10, 40, 3, 0, 0.02, "bins"),
}


def scaled(cfg: int, n: int | None = None, **over) -> SynthConfig:
    """A config with ``n`` (and any other field) overridden, e.g. for tests."""
    base = CONFIGS[cfg]
    d = dict(base.__dict__)
    if n is not None:
        d["n"] = int(n)
    d.update(over)
    d["G_u"] = min(d["G_u"], d["G"])
    return SynthConfig(**d)


def make_coords(c: SynthConfig, rng) -> np.ndarray:
    n = c.n
    if c.coords == "hex":
        side = int(np.ceil(np.sqrt(n)))
        ii, jj = np.divmod(np.arange(n), side)
        x = (jj + 0.5 * (ii & 1)) * 100.0
        y = ii * (100.0 * np.sqrt(3.0) / 2.0)
        xy = np.stack([x, y], 1) + rng.normal(0.0, 1.0, (n, 2))
    elif c.coords == "uniform":
        xy = rng.random((n, 2)) * (10.0 * np.sqrt(n))
    elif c.coords == "bins":
        side = int(np.ceil(np.sqrt(n)))
        ii, jj = np.divmod(np.arange(n), side)
        xy = np.stack([jj, ii], 1).astype(np.float64) + rng.uniform(-0.01, 0.01, (n, 2))
    else:
        raise ValueError(c.coords)
    return np.ascontiguousarray(xy, dtype=np.float64)


def make_labels(c: SynthConfig, xy: np.ndarray, rng) -> np.ndarray:
    """Spatially coherent labels: nearest of 10*C random Voronoi seeds."""
    from scipy.spatial import cKDTree
    lo, hi = xy.min(0), xy.max(0)
    seeds = lo + rng.random((10 * c.C, 2)) * (hi - lo)
    seed_type = rng.integers(0, c.C, 10 * c.C)
    seed_type[: c.C] = np.arange(c.C)          # every type occurs
    _, nearest = cKDTree(seeds).query(xy)
    return seed_type[nearest].astype(np.int32)


def make_expression(c: SynthConfig, labels: np.ndarray, rng, chunk: int = 65536) -> sp.csr_matrix:
    """CSR float32 n x G, values log1p(1+NegBin), density modulated by type."""
    base = rng.beta(2.0, 2.0 * (1.0 - c.density) / c.density, c.G)          # mean == density
    mod = rng.gamma(2.0, 0.5, (c.C, c.G))                                    # mean 1
    dens = np.clip(base[None, :] * mod, 0.0, 0.95).astype(np.float32)       # C x G
    blocks = []
    for s in range(0, c.n, chunk):
        lab = labels[s:s + chunk]
        hit = rng.random((len(lab), c.G), dtype=np.float32) < dens[lab]
        r, g = np.nonzero(hit)
        v = np.log1p(1.0 + rng.negative_binomial(2, 0.4, r.size)).astype(np.float32)
        blocks.append(sp.csr_matrix((v, (r, g)), shape=(len(lab), c.G)))
    X = sp.vstack(blocks, format="csr")
    X.sort_indices()
    return X


_FAST_VALUES = np.log1p(1.0 + np.arange(16, dtype=np.float64)).astype(np.float32)


def make_expression_fast(c: SynthConfig, rng) -> sp.csr_matrix:
    """A cheaper generator for the full-size parity tests of configs 3-5 (tens of seconds instead of minutes): the
    non-zeros of the n x G matrix are placed by geometric gaps over the row-major index space (uniform density
    ``c.density``, no cell-type modulation), values from a 16-entry table of log1p counts.  Canonical CSR."""
    total = c.n * c.G
    m = int(total * c.density * 1.02) + 4096
    pos = np.cumsum(rng.geometric(c.density, m), dtype=np.int64) - 1
    pos = pos[: np.searchsorted(pos, total)]
    rows, cols = np.divmod(pos, c.G)
    del pos
    indptr = np.zeros(c.n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=c.n), out=indptr[1:])
    del rows
    vals = _FAST_VALUES[rng.integers(0, 16, len(cols), dtype=np.uint8)]
    X = sp.csr_matrix((vals, cols.astype(np.int32), indptr), shape=(c.n, c.G))
    X.has_sorted_indices = True
    return X


def make_lr_table(c: SynthConfig, rng) -> pd.DataFrame:
    """P distinct (ligand, receptor) rows over the first G_u genes.

    Multi-subunit complexes are pre-expanded to one row per (ligand, subunit),
    the way the reference's tutorial formats CellChatDB
    (docs/notebooks/04_LARIS_tutorial.html:743,767)."""
    gu = c.G_u
    half = gu // 2
    seen, lig, rec = set(), [], []
    while len(lig) < c.P:
        l = int(rng.integers(0, half))
        n_sub = 1
        if c.complex_frac > 0 and rng.random() < c.complex_frac:
            n_sub = int(rng.integers(2, 4))
        subs = rng.choice(np.arange(half, gu), n_sub, replace=False)
        for r in subs:
            if (l, int(r)) not in seen and len(lig) < c.P:
                seen.add((l, int(r)))
                lig.append(l)
                rec.append(int(r))
    names = gene_names(c.G)
    return pd.DataFrame({"ligand": names[lig], "receptor": names[rec]})


def gene_names(G: int) -> np.ndarray:
    return np.array([f"g{g:05d}" for g in range(G)], dtype=object)


def make_dataset(cfg: int, n: int | None = None, seed: int | None = None, fast: bool = False, **over):
    """Return ``(adata, lr_df, config)`` for BASELINE config ``cfg``.  ``fast=True`` swaps the expression generator
    for ``make_expression_fast`` (different bytes, same shape / density; used by the full-size parity tests)."""
    c = scaled(cfg, n, **over)
    rng = np.random.default_rng(1000 + cfg if seed is None else seed)
    xy = make_coords(c, rng)
    labels = make_labels(c, xy, rng)
    X = make_expression_fast(c, rng) if fast else make_expression(c, labels, rng)
    lr_df = make_lr_table(c, rng)
    cats = [f"ct{t:02d}" for t in range(c.C)]
    obs = pd.DataFrame(
        {"CellTypes": pd.Categorical.from_codes(labels, categories=cats)},
        index=pd.Index([f"c{i}" for i in range(c.n)]),
    )
    var = pd.DataFrame(index=pd.Index(gene_names(c.G)))
    adata = make_anndata(X, obs=obs, var=var, obsm={"X_spatial": xy})
    return adata, lr_df, c
