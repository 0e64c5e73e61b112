"""Host-side stand-ins for the three ``cosg`` helpers the reference calls.

The reference imports the un-vendored package ``cosg>=1.0.3``
(reference pyproject.toml:39) and calls ``cosg.cosg`` (laris/tools/_utils.py:580-588),
``cosg.indexByGene`` (:590-595, :971-974) and ``cosg.iqrLogNormalize`` (:596, :975).
The package is not installable here and its source is not in the reference
tree, so:

* when ``import cosg`` works, the real ``indexByGene`` / ``iqrLogNormalize`` are used;
* otherwise the functions below are used.  ``index_by_gene`` is a plain pivot
  of the ranked ``names@@group`` / ``scores@@group`` frame and is determined by
  the frame layout the reference itself builds (laris/tools/_utils.py:945-953).
  ``iqr_log_normalize`` is a LABELLED STAND-IN (monotone per-column rescale
  to [0, 1]); it is **excluded from every parity claim** (SURVEY.md §8c,
  DESIGN.md "cosg"), and ``runLARIS(by_celltype=True)`` warns loudly whenever
  it is in use.

The COSG score itself (cosine to the one-hot group indicator + the
``cos^3/((1-mu)cos^2 + mu*sum cos^2)`` regularisation, which the reference
clones in laris/tools/_utils.py:861-880) runs on the device
(``laris_onehot_contract_rows`` for genes, ``laris_celltype_pair_contract`` +
``laris_pair_cosine_regularise`` for the neighbourhood table, include/laris_b200.h);
only table reshaping lives here.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


class StringTaker:
    """``values[codes]`` as DataFrame columns, many times for one ``values``.  With pandas' Arrow-backed default string
    dtype a column is built by decoding an Arrow dictionary (a few unique names, up to millions of rows) instead of
    converting an object array element by element - same dtype, same values, ~6x faster for the 3.2 M-row cell-type
    table and the 3200-column ranked frame of config 3.  Anything else (non-string values, no pyarrow, older pandas)
    falls back to the plain gather."""

    def __init__(self, values):
        self.values = np.asarray(values, dtype=object)
        self.dtype = self.dictionary = None
        try:
            if len(self.values) and all(isinstance(v, str) for v in self.values):
                probe = pd.Series(self.values[:1]).dtype
                if isinstance(probe, pd.StringDtype) and getattr(probe, "storage", None) == "pyarrow":
                    import pyarrow as pa
                    self._pa = pa
                    self.dictionary = pa.array(list(self.values), type=pa.large_string())
                    self.dtype = probe
        except Exception:        # noqa: BLE001 - any pyarrow / pandas version quirk: plain gather
            self.dtype = self.dictionary = None

    def take(self, codes):
        codes = np.asarray(codes)
        if self.dictionary is not None:
            try:
                return pd.array(self._decode(codes), dtype=self.dtype)
            except Exception:    # noqa: BLE001
                pass
        return self.values[codes]

    def _decode(self, codes):
        """Arrow string array ``dictionary[codes]`` (the GIL is released while it is built)."""
        pa = self._pa
        if codes.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
            codes = codes.astype(np.int64)
        return pa.DictionaryArray.from_arrays(pa.array(codes), self.dictionary).dictionary_decode()


def take_strings(values, codes):
    return StringTaker(values).take(codes)


_TAKE_CHUNK = 1 << 19


def _pool():
    """Worker threads for the string columns, or None on a single-core host."""
    from . import _threads
    return _threads.pool()


def take_many_async(jobs):
    """Start ``[taker.take(codes) for taker, codes in jobs]`` on worker threads and return ``collect()``.  Every column is
    decoded in chunks of 2^19 rows side by side (decoding an Arrow dictionary releases the GIL) and handed to pandas as one
    chunked Arrow array, so the five string columns of the 3.2 M-row cell-type table cost a few ms of wall time instead of
    5 x 17 ms - and the caller can do other work (or wait for the device) in between."""
    pool = _pool()
    small = not jobs or max(len(c) for _, c in jobs) < 200_000
    if pool is None or small or any(t.dictionary is None for t, _ in jobs):
        done = [t.take(c) for t, c in jobs]
        return lambda: done
    pending = []
    for t, c in jobs:
        c = np.asarray(c)
        pending.append((t, c, [pool.submit(t._decode, c[s:s + _TAKE_CHUNK]) for s in range(0, len(c), _TAKE_CHUNK)]))

    def collect():
        out = []
        for t, c, futs in pending:
            try:
                out.append(pd.array(t._pa.chunked_array([f.result() for f in futs]), dtype=t.dtype))
            except Exception:    # noqa: BLE001 - any pyarrow / pandas quirk: the plain gather
                out.append(t.values[c])
        return out
    return collect


def take_many(jobs):
    return take_many_async(jobs)()


_REAL_COSG = [False, None]          # [looked up, module]: a failing ``import cosg`` walks sys.path every time it is tried


def real_cosg():
    """The installed ``cosg`` package, or None (looked up once per process; a stub module put into ``sys.modules`` by the
    oracle is not the real thing)."""
    import sys
    mod = sys.modules.get("cosg")
    if mod is not None and getattr(mod, "__laris_stub__", False):
        return None
    if not _REAL_COSG[0]:
        _REAL_COSG[0] = True
        try:
            import cosg  # type: ignore
            if hasattr(cosg, "indexByGene") and hasattr(cosg, "iqrLogNormalize") and not getattr(cosg, "__laris_stub__", False):
                _REAL_COSG[1] = cosg
        except ImportError:
            pass
    return _REAL_COSG[1]


def ranked_frame(names, scores, groups, delimiter="@@"):
    """Per-group descending ranking as the reference stores it in ``uns``.

    ``scores`` is [n_names, n_groups].  Columns are ``names{d}{group}`` and
    ``scores{d}{group}``, rows are ranks (laris/tools/_utils.py:905-953).
    Returns (frame, order) with ``order[:, g]`` the name index at each rank.
    """
    names = np.asarray(names, dtype=object)
    scores = np.asarray(scores)
    # reference: argpartition + argsort[::-1] of the full column (:139-143);
    # descending, ties in reversed-ascending order
    order = np.argsort(scores, axis=0, kind="stable")[::-1]
    cols = {}
    taker = StringTaker(names)
    for g, grp in enumerate(groups):
        cols[f"names{delimiter}{grp}"] = taker.take(order[:, g])
    for g, grp in enumerate(groups):
        cols[f"scores{delimiter}{grp}"] = scores[order[:, g], g]
    return pd.DataFrame(cols), order


def index_by_gene(df, gene_key="names", score_key="scores", set_nan_to_zero=False,
                  convert_negative_one_to_zero=True, column_delimiter="@@"):
    """Pivot a ranked frame back to a names x groups score table."""
    cosg = real_cosg()
    if cosg is not None:
        return cosg.indexByGene(df, gene_key=gene_key, score_key=score_key,
                                set_nan_to_zero=set_nan_to_zero,
                                convert_negative_one_to_zero=convert_negative_one_to_zero,
                                column_delimiter=column_delimiter)
    prefix = f"{gene_key}{column_delimiter}"
    groups = [c[len(prefix):] for c in df.columns if str(c).startswith(prefix)]
    ng = len(groups)
    nm = np.concatenate([df[f"{prefix}{g}"].to_numpy(dtype=object) for g in groups]) if ng else np.empty(0, object)
    codes, all_names = pd.factorize(nm)                  # first-appearance order, one hash pass for all groups
    vals = np.full((len(all_names), ng), np.nan)
    rows = len(df)
    for j, g in enumerate(groups):
        vals[codes[j * rows:(j + 1) * rows], j] = df[f"{score_key}{column_delimiter}{g}"].to_numpy(dtype=np.float64)
    out = pd.DataFrame(vals, index=pd.Index(all_names, dtype=object), columns=groups)
    if convert_negative_one_to_zero:
        out = out.mask(out == -1, 0.0)
    if set_nan_to_zero:
        out = out.fillna(0.0)
    return out


def index_by_gene_of_ranking(frame, names, scores, groups, order, set_nan_to_zero=False,
                             convert_negative_one_to_zero=True, column_delimiter="@@"):
    """``index_by_gene(frame)`` for a frame that ``ranked_frame(names, scores, groups)`` just built: with the real
    ``cosg`` the frame goes through ``cosg.indexByGene``; otherwise the pivot of a complete ranking is the score matrix
    itself (rows in first-appearance order = the ranking of the first group), so the 2 * len(groups) string / float
    columns are not parsed back."""
    cosg = real_cosg()
    if cosg is not None:
        return cosg.indexByGene(frame, gene_key="names", score_key="scores", set_nan_to_zero=set_nan_to_zero,
                                convert_negative_one_to_zero=convert_negative_one_to_zero,
                                column_delimiter=column_delimiter)
    names = np.asarray(names, dtype=object)
    rows = order[:, 0] if len(groups) else np.arange(len(names))
    vals = np.asarray(scores, dtype=np.float64)[rows]
    if convert_negative_one_to_zero:
        vals = np.where(vals == -1, 0.0, vals)
    if set_nan_to_zero:
        vals = np.nan_to_num(vals, nan=0.0)
    return pd.DataFrame(vals, index=pd.Index(names[rows], dtype=object), columns=list(groups))


def iqr_log_normalize(df, q_upper=0.95, q_lower=0.75):
    """STAND-IN for ``cosg.iqrLogNormalize`` - parity unpinned (the package is not vendored by the reference and not
    installable here; the reference only documents that the result is a specificity score in [0, 1],
    laris/tools/_utils.py:556-561).

    Per column: scale = Q_upper - Q_lower of the positive entries (their maximum when that difference is not
    positive), y = log1p(max(x, 0) / scale), result = y / max(y): monotone, zeros stay zero, the most specific
    entry of every column is 1.  ``laris_column_iqr_log`` (include/laris_b200.h, T2) is the same arithmetic on the
    device.  With the real ``cosg`` installed this function is never used.
    """
    cosg = real_cosg()
    if cosg is not None:
        return cosg.iqrLogNormalize(df)
    return pd.DataFrame(iqr_log_normalize_array(df.to_numpy(dtype=np.float64), q_upper, q_lower), index=df.index,
                        columns=df.columns)


def iqr_log_normalize_array(v, q_upper=0.95, q_lower=0.75):
    """The stand-in above on a plain [rows, columns] fp64 array, all columns at once: one sort per column, the two
    quantiles of the positive entries by numpy's 'linear' rule (same ``_lerp`` arithmetic as ``numpy.quantile``)."""
    v = np.asarray(v, dtype=np.float64)
    rows, cols = v.shape
    out = np.zeros_like(v)
    if rows == 0 or cols == 0:
        return out
    pos = v > 0
    m = pos.sum(axis=0)                                             # positive entries per column
    srt = np.sort(np.where(pos, v, np.inf), axis=0)                  # positives ascending, then +inf
    ar = np.arange(cols)
    live = m > 0
    mm = np.maximum(m, 1)

    def quantile(q):
        h = (mm - 1) * q
        lo = np.floor(h).astype(np.int64)
        hi = np.minimum(lo + 1, mm - 1)
        a, b, t = srt[lo, ar], srt[hi, ar], h - lo
        d = b - a
        return np.where(t >= 0.5, b - d * (1 - t), a + d * t)

    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        mx = np.where(live, srt[mm - 1, ar], 1.0)
        scale = quantile(q_upper) - quantile(q_lower)
        scale = np.where(scale > 0, scale, mx)
        y = np.log1p(np.maximum(v, 0.0) / scale[None, :]) / np.log1p(mx / scale)[None, :]
    out[:, live] = y[:, live]
    return out


_WARNED = False


def warn_if_stand_in():
    """One loud warning per process when the cell-type step runs without the real ``cosg``: its interaction scores
    then go through the stand-in normalisation above and DIFFER from the reference's (ranking within a column is
    preserved, values are not)."""
    global _WARNED
    if _WARNED or real_cosg() is not None:
        return
    _WARNED = True
    import warnings
    warnings.warn(
        "laris_b200: the 'cosg' package (cosg>=1.0.3) is not importable; cosg.indexByGene / cosg.iqrLogNormalize are "
        "replaced by laris_b200.cosg_compat stand-ins whose parity with cosg is UNPINNED. Cell-type interaction scores, "
        "their p-values and FDR will differ from genecell/LARIS until cosg is installed.", RuntimeWarning, stacklevel=3)
