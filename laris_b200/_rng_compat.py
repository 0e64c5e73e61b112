"""Host replay of the reference's numpy legacy (MT19937) random stream.

``rng="numpy"`` reproduces the reference's draws exactly: it seeds
``np.random.seed(random_seed)``, draws the per-repeat seeds with
``choice(1902772, n_repeats, replace=False)`` (reference laris/tools/_utils.py:497-499),
builds each random graph with ``seed(s); choice(arange(n), n*k); shuffle(weights)``
(:437-442), and the p-value draws continue from the state left by the LAST
repeat (the reference never reseeds before ``randint`` at :1546).  Only index
generation happens here; all arithmetic on the draws runs on the device.

A private ``RandomState`` is used; the final state is copied into numpy's
global generator when ``sync_global=True`` so callers that relied on the
reference's side effect on ``np.random`` see the same stream afterwards.
"""
from __future__ import annotations

import numpy as np


def repeat_seeds(random_seed: int, n_repeats: int):
    return np.random.RandomState(int(random_seed)).choice(1902772, size=n_repeats, replace=False)


def null_graph(n: int, k: int, seed_r: int):
    """(cols [n*k] int64, perm [n*k] int64, RandomState) with shuffled == weights[perm]."""
    rs = np.random.RandomState(int(seed_r))
    cols = rs.choice(np.arange(n), n * k, replace=True)
    perm = np.arange(n * k)
    rs.shuffle(perm)
    return cols.astype(np.int64), perm.astype(np.int64), rs


def pvalue_draws(rs: np.random.RandomState, n_rows: int, n_perm: int, nb: int):
    """``np.random.randint(0, nb, size=(rows, N))`` on the continued stream (:1546)."""
    return rs.randint(0, nb, size=(n_rows, n_perm))


def sync_global(rs: np.random.RandomState):
    np.random.set_state(rs.get_state())
