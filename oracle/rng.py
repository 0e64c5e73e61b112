"""CPU twin of the device counter-based RNG.  TEST INFRASTRUCTURE ONLY.

The reference draws from numpy's global legacy MT19937 stream
(reference laris/tools/_utils.py:437-442, :497-499, :1546).  The device path
offers two modes (DESIGN.md "RNG"):

* ``rng="numpy"``  - the host replays the reference's legacy stream and uploads
  the draws; bit-identical to the reference.  Helpers: :func:`numpy_null_graph`,
  :func:`numpy_pvalue_draws`.
* ``rng="philox"`` - Philox4x32-10 keyed by ``random_seed`` with the counter set
  to the global (repeat, edge) / (row, permutation) id, so results do not
  depend on how work is sharded over GPUs.  This file restates that generator
  in numpy so the oracle can be fed exactly the device's draws.

Stream layout (must match laris_b200/csrc/rng.cuh):
  counter = (c0, c1, c2, stream)   key = (seed & 0xffffffff, seed >> 32)
  STREAM_COL  = 1 : c0,c1 = edge id lo/hi, c2 = repeat; col = mulhi64(o0|o1<<32, n)
  STREAM_PERM = 2 : c0 = round, c1 = 0, c2 = repeat; Feistel round key = o0
  STREAM_PVAL = 3 : c0 = t // 4, c1,c2 = row id lo/hi; draw[t%4] = mulhi32(o[t%4], nb)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
STREAM_COL, STREAM_PERM, STREAM_PVAL = 1, 2, 3
FEISTEL_ROUNDS = 6


def philox4x32(c0, c1, c2, c3, seed: int):
    """Vectorised Philox4x32-10.  Counters broadcast; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & MASK32
                                           for c in (c0, c1, c2, c3)])
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0            # < 2^64, exact in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def mulhi64(a, b: int):
    """floor(a * b / 2^64) for uint64 array ``a`` and python int ``b`` < 2^64."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.uint64(b)
    a_lo, a_hi = a & MASK32, a >> np.uint64(32)
    b_lo, b_hi = b & MASK32, b >> np.uint64(32)
    ll = a_lo * b_lo
    lh = a_lo * b_hi
    hl = a_hi * b_lo
    hh = a_hi * b_hi
    mid = (ll >> np.uint64(32)) + (lh & MASK32) + (hl & MASK32)
    return hh + (lh >> np.uint64(32)) + (hl >> np.uint64(32)) + (mid >> np.uint64(32))


def philox_null_cols(n: int, n_edges: int, repeat: int, seed: int, start: int = 0):
    """Random neighbour column of every edge of null graph ``repeat``."""
    e = np.arange(start, start + n_edges, dtype=np.uint64)
    o0, o1, _, _ = philox4x32(e & MASK32, e >> np.uint64(32), repeat, STREAM_COL, seed)
    u = o0.astype(np.uint64) | (o1.astype(np.uint64) << np.uint64(32))
    return mulhi64(u, n).astype(np.int64)


def _fmix32(x):
    x = np.asarray(x, dtype=np.uint64) & MASK32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & MASK32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & MASK32
    x ^= x >> np.uint64(16)
    return x


def feistel_keys(repeat: int, seed: int):
    rd = np.arange(FEISTEL_ROUNDS, dtype=np.uint64)
    o0, _, _, _ = philox4x32(rd, 0, repeat, STREAM_PERM, seed)
    return o0.astype(np.uint64)


def feistel_half_bits(m: int) -> int:
    b = max(2, int(m - 1).bit_length())
    return (b + 1) // 2


def feistel_permutation(m: int, repeat: int, seed: int, idx=None):
    """Bijection of [0, m): balanced Feistel network + cycle walking."""
    h = np.uint64(feistel_half_bits(m))
    mask = (np.uint64(1) << h) - np.uint64(1)
    keys = feistel_keys(repeat, seed)
    x = np.arange(m, dtype=np.uint64) if idx is None else np.asarray(idx, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        xs = x[todo]
        L, R = xs >> h, xs & mask
        for rd in range(FEISTEL_ROUNDS):
            f = _fmix32(((R * np.uint64(0x9E3779B1)) & MASK32) + keys[rd]) & mask
            L, R = R, L ^ f
        xs = (L << h) | R
        x[todo] = xs
        todo[todo] = xs >= np.uint64(m)
    return x.astype(np.int64)


def philox_pvalue_draws(row_ids, n_perm: int, nb: int, seed: int):
    """draws[row, t] in [0, nb) for global row ids (row = group * P + pair).

    For nb <= 64 every 32-bit Philox word w yields three draws, the first three base-nb digits of w / 2^32
    (d = (w * nb) >> 32; w = (w * nb) mod 2^32), so one Philox call covers 12 permutations; else one draw per word."""
    digits = 3 if nb <= 64 else 1
    per_call = 4 * digits
    row = np.asarray(row_ids, dtype=np.uint64)[:, None]
    q = np.arange((n_perm + per_call - 1) // per_call, dtype=np.uint64)[None, :]
    outs = philox4x32(q, row & MASK32, row >> np.uint64(32), STREAM_PVAL, seed)
    w = np.stack(outs, axis=-1).astype(np.uint64)                      # [rows, calls, 4]
    ds = []
    for _ in range(digits):
        x = w * np.uint64(nb)
        ds.append(x >> np.uint64(32))
        w = x & MASK32
    d = np.stack(ds, axis=-1).reshape(row.shape[0], -1)[:, :n_perm]    # [rows, calls * 4 * digits]
    return d.astype(np.int64)


# ---------------------------------------------------------------------------
# numpy-legacy replay (reference-compatible stream)
# ---------------------------------------------------------------------------

def numpy_repeat_seeds(random_seed: int, n_repeats: int):
    """reference _utils.py:497-499."""
    rs = np.random.RandomState(random_seed)
    return rs.choice(1902772, size=n_repeats, replace=False)


def numpy_null_graph(n: int, k: int, seed_r: int):
    """(cols, perm) of one random graph: reference _utils.py:436-442.

    ``shuffle(x)`` after ``seed(s); choice(arange(n), n*k)`` equals
    ``x[perm]`` for the permutation returned here (SURVEY.md §8c probe).
    Returns the RandomState so a caller can continue the stream (the
    reference's p-value draws continue from the last repeat's state).
    """
    rs = np.random.RandomState(int(seed_r))
    cols = rs.choice(np.arange(n), n * k, replace=True)
    perm = np.arange(n * k)
    rs.shuffle(perm)
    return cols.astype(np.int64), perm.astype(np.int64), rs
