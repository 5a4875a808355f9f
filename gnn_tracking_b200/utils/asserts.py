"""Shape checks at the boundary (reference utils/asserts.py:4-7)."""
from torch import Tensor


def assert_feat_dim(feat_vec: Tensor, dim: int) -> None:
    assert feat_vec.shape[1] == dim, f"expected feature width {dim}, got {tuple(feat_vec.shape)}"
