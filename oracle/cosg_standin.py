"""TEST INFRASTRUCTURE - oracle-side stand-ins for the three helpers the reference takes from the un-vendored
package ``cosg>=1.0.3`` (reference pyproject.toml:39; call sites laris/tools/_utils.py:580-596, :971-975).

PARITY UNPINNED: the package is absent here and the reference ships neither its source nor a test of it, so these
are restatements of documented behaviour only:

* ``ranked_frame`` / ``index_by_gene``: the ``names@@group`` / ``scores@@group`` layout is fixed by the frame the
  reference itself builds at laris/tools/_utils.py:905-953 and reads back at :971-974 - a pivot.
* ``iqr_log_normalize``: the reference documents only "normalized specificity scores [0, 1]" (:556-561).  The stand-in
  is: per column, scale = Q95 - Q75 of the positive entries (their maximum if that is not positive),
  y = log1p(max(x, 0) / scale), result = y / max(y).

They are written independently of ``laris_b200.cosg_compat`` (the product's own stand-ins) so that the golden
fixtures generated through them check the product instead of mirroring it.  Nothing under ``laris_b200/`` imports
this file.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def ranked_frame(names, scores, groups, delimiter="@@"):
    """Every group's names and scores by descending score (ties: reversed ascending stable order, which is what
    ``argpartition`` + ``argsort[::-1]`` of the full column gives, laris/tools/_utils.py:139-143)."""
    names = np.asarray(names, dtype=object)
    scores = np.asarray(scores, dtype=np.float64)
    data = {}
    orders = []
    for j, g in enumerate(groups):
        o = np.argsort(scores[:, j], kind="stable")[::-1]
        orders.append(o)
        data[f"names{delimiter}{g}"] = names[o]
    for j, g in enumerate(groups):
        data[f"scores{delimiter}{g}"] = scores[orders[j], j]
    return pd.DataFrame(data)


def index_by_gene(df, gene_key="names", score_key="scores", set_nan_to_zero=False, convert_negative_one_to_zero=True,
                  column_delimiter="@@"):
    """names x groups score table of a ranked frame; rows in first-appearance order."""
    prefix = gene_key + column_delimiter
    groups = [str(c)[len(prefix):] for c in df.columns if str(c).startswith(prefix)]
    seen = {}
    for g in groups:
        for nm in df[prefix + g]:
            seen.setdefault(nm, len(seen))
    out = np.full((len(seen), len(groups)), np.nan)
    for j, g in enumerate(groups):
        rows = [seen[nm] for nm in df[prefix + g]]
        out[rows, j] = df[score_key + column_delimiter + g].to_numpy(dtype=np.float64)
    if convert_negative_one_to_zero:
        out[out == -1] = 0.0
    if set_nan_to_zero:
        out[np.isnan(out)] = 0.0
    return pd.DataFrame(out, index=pd.Index(list(seen), dtype=object), columns=groups)


def iqr_log_normalize(df, q_upper=0.95, q_lower=0.75):
    v = np.asarray(df, dtype=np.float64)
    out = np.zeros_like(v)
    for j in range(v.shape[1]):
        x = v[:, j]
        pos = np.sort(x[x > 0])
        if len(pos) == 0:
            continue

        def quantile(q):                     # numpy 'linear' definition, spelled out
            h = (len(pos) - 1) * q
            lo = int(np.floor(h))
            hi = min(lo + 1, len(pos) - 1)
            return pos[lo] + (pos[hi] - pos[lo]) * (h - lo)

        scale = quantile(q_upper) - quantile(q_lower)
        if not scale > 0:
            scale = pos[-1]
        out[:, j] = np.log1p(np.maximum(x, 0.0) / scale) / np.log1p(pos[-1] / scale)
    if isinstance(df, pd.DataFrame):
        return pd.DataFrame(out, index=df.index, columns=df.columns)
    return out
