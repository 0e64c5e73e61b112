"""laris_b200.pp - the three helpers the reference re-exports for backward
compatibility (reference laris/preprocessing/__init__.py:28-38)."""
from ..tools._utils import _pairwise_row_multiply, _rowwise_cosine_similarity, _select_top_n

__all__ = ["_rowwise_cosine_similarity", "_select_top_n", "_pairwise_row_multiply"]
