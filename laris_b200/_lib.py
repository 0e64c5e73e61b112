"""ctypes binding of ``liblaris_b200.so`` (C ABI declared in include/laris_b200.h).

There is no CPU fallback: if the shared library has not been built this module
raises at import time, and every wrapper raises ``RuntimeError`` with the
library's own error text when a call returns non-zero.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblaris_b200.so")

if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). laris_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_p, _i, _i64, _u64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double

# name -> argtypes; must list every LARIS_API symbol of include/laris_b200.h
SIGNATURES = {
    "laris_abi_version": [],
    "laris_knn_graph": [_p, _i64, _i, _i, _i, _p, _p, _p, _p],
    "laris_decay_weights": [_p, _i64, _d, _p, _p, _p, _p],
    "laris_radius_graph_count": [_p, _i64, _i, _d, _i, _p, _p, _p],
    "laris_radius_graph_fill": [_p, _i64, _i, _d, _i, _p, _p, _p, _p],
    "laris_csr_pad_rows": [_p, _p, _p, _i64, _i, _p, _p, _p],
    "laris_csr_select_count": [_p, _p, _p, _i64, _p, _p, _p],
    "laris_csr_select_fill": [_p, _p, _p, _i64, _p, _p, _p, _p],
    "laris_csr_select_rows": [_p, _p, _p, _i64, _p, _p, _p, _p],
    "laris_exclusive_scan_i64": [_p, _i64, _p, _p],
    "laris_diffuse_lr_score": [_p, _p, _i64, _i, _p, _p, _i, _p, _p, _p, _i, _p, _i64, _p, _p],
    "laris_diffuse_lr_score_complex": [_p, _p, _i64, _i, _p, _p, _i, _p, _p, _p, _i, _p, _i64, _p, _p, _p, _p, _i, _i, _p],
    "laris_diffuse_lr_score_rows": [_p, _p, _p, _i64, _i, _p, _p, _i, _p, _p, _p, _i, _p, _i64, _p, _p, _p, _p, _i, _i, _p],
    "laris_diffuse_lr_score_dense_complex": [_p, _i64, _p, _i64, _i, _p, _p, _i, _p, _p, _p, _i, _p, _i64, _p, _p, _p, _p,
                                             _i, _i, _p],
    "laris_csr_select_dense": [_p, _p, _p, _i64, _p, _p, _i64, _p, _p],
    "laris_diffuse_lr_score_dense": [_p, _i64, _p, _i64, _i, _p, _p, _i, _p, _p, _p, _i, _p, _i64, _p, _p],
    "laris_lr_column_layout": [_p, _p, _i, _i, _p, _p],
    "laris_csr_compact": [_p, _i64, _i64, _i, _p, _p, _p, _p],
    "laris_csr_to_dense": [_p, _p, _p, _i64, _i, _p, _i64, _p],
    "laris_spatial_stats": [_p, _i64, _i64, _i, _p, _p, _i, _i, _p, _p, _p],
    "laris_spatial_stats_batched": [_p, _i64, _i64, _i, _p, _p, _i, _p, _p, _p, _i, _p, _p],
    "laris_spatial_stats_sparse_max_ld": [],
    "laris_scores_pack": [_p, _i64, _i64, _i, _p, _p, _i64, _p],
    "laris_spatial_stats_sparse": [_p, _i64, _i64, _i, _p, _p, _i64, _p, _p, _i, _p, _p, _p, _i, _p, _p],
    "laris_null_graph_philox": [_i64, _i, _p, _i, _u64, _p, _p, _p],
    "laris_gather_f32": [_p, _p, _i64, _p, _p],
    "laris_neighborhood_composition": [_p, _p, _p, _i64, _i, _i, _p, _p],
    "laris_celltype_pair_contract": [_p, _i64, _i64, _i, _p, _i, _p, _p, _p, _i, _p],
    "laris_celltype_pair_keys": [_p, _i64, _i, _p, _p],
    "laris_celltype_pair_contract_sparse": [_p, _i64, _i64, _i, _p, _i, _p, _i64, _p, _p],
    "laris_pair_profile_norm": [_p, _i64, _i, _p, _p],
    "laris_permute_labels": [_p, _i64, _i, _u64, _p, _p],
    "laris_count_exceed": [_p, _p, _i64, _p, _p],
    "laris_celltype_nonzero_count": [_p, _i64, _i64, _i, _p, _i, _p, _p],
    "laris_celltype_nonzero_count_packed": [_p, _p, _i64, _i64, _i, _p, _i, _p, _p],
    "laris_onehot_contract_csr": [_p, _p, _i64, _i, _p, _i, _p, _p, _p, _p],
    "laris_onehot_contract_rows": [_p, _p, _p, _i64, _i, _p, _i, _p, _p, _p, _p],
    "laris_pvalue_counts": [_p, _i64, _i, _i, _p, _i, _p, _p, _i64, _i, _i, _i, _i, _u64, _p, _p, _p, _p, _p],
    "laris_bh_fdr": [_p, _p, _i, _i, _p, _p],
    "laris_pair_cosine_regularise": [_p, _p, _p, _i, _i, _d, _p, _p, _p],
    "laris_column_iqr_log": [_p, _i64, _i64, _d, _d, _p, _p],
    "laris_interaction_scores": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p],
    "laris_topk_mean": [_p, _i64, _i, _p, _p],
    "laris_finalize_scores": [_p, _p, _p, _p, _i, _i, _i, _d, _p, _p, _p, _p],
    "laris_pvalues_from_counts": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _d, _p, _p, _p, _p],
    "laris_gather_result_rows": [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p],
    "laris_result_row_order": [_p, _p, _i, _i, _p, _p, _p],
    "laris_result_rows_gather": [_p, _i64, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "laris_row_qc": [_p, _i64, _i64, _i, _p, _i, _p, _p, _p, _p],
    "laris_rowwise_cosine": [_p, _p, _i64, _i64, _i64, _i64, _p, _p],
    "laris_pairwise_row_multiply": [_p, _i, _i64, _p, _p],
}

for _name, _args in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = C.c_int
lib.laris_launch_count.argtypes = []
lib.laris_launch_count.restype = C.c_int64
lib.laris_last_error.argtypes = []
lib.laris_last_error.restype = C.c_char_p


class LarisError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        raise LarisError(f"{what} failed (rc={rc}): {lib.laris_last_error().decode(errors='replace')}")
