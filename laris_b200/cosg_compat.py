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
  ``iqr_log_normalize`` is a LABELLED PLACEHOLDER (monotone per-column
  rescale); it is **excluded from every parity claim** (SURVEY.md §8c,
  DESIGN.md "cosg").

The COSG score itself (cosine to the one-hot group indicator + the
``cos^3/((1-mu)cos^2 + mu*sum cos^2)`` regularisation, which the reference
clones in laris/tools/_utils.py:861-880) runs on the device
(``laris_celltype_cosine`` in include/laris_b200.h); only table reshaping
lives here.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def real_cosg():
    try:
        import cosg  # type: ignore
        if (hasattr(cosg, "indexByGene") and hasattr(cosg, "iqrLogNormalize")
                and not getattr(cosg, "__laris_stub__", False)):
            return cosg
    except ImportError:
        pass
    return None


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
    for g, grp in enumerate(groups):
        cols[f"names{delimiter}{grp}"] = names[order[:, g]]
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


def iqr_log_normalize(df, q_upper=0.95, q_lower=0.75):
    """PLACEHOLDER for ``cosg.iqrLogNormalize`` - parity unpinned.

    Per column: scale = Q_upper - Q_lower of the positive entries (max if that
    is 0), result = log1p(max(x,0)/scale).  Monotone, keeps zeros at zero.
    """
    cosg = real_cosg()
    if cosg is not None:
        return cosg.iqrLogNormalize(df)
    v = df.to_numpy(dtype=np.float64)
    out = np.zeros_like(v)
    for j in range(v.shape[1]):
        col = v[:, j]
        pos = col[col > 0]
        if pos.size == 0:
            continue
        scale = np.quantile(pos, q_upper) - np.quantile(pos, q_lower)
        if not scale > 0:
            scale = pos.max()
        out[:, j] = np.log1p(np.maximum(col, 0.0) / scale)
    return pd.DataFrame(out, index=df.index, columns=df.columns)
