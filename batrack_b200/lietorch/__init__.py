import torch

from .groups import SE3  # noqa: F401


def cat(group_list, dim):
    """Concatenate groups along a dimension (main/backend/lietorch/groups.py:314-317)."""
    return group_list[0].__class__(torch.cat([X.data for X in group_list], dim=dim))


def stack(group_list, dim):
    """Stack groups along a new dimension (main/backend/lietorch/groups.py:319-322)."""
    return group_list[0].__class__(torch.stack([X.data for X in group_list], dim=dim))


__all__ = ["SE3", "cat", "stack"]
