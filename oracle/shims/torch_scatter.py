"""ORACLE SHIM (test infrastructure): stand-in for torch-scatter 2.1.2's `scatter_sum`
(reference requirements.txt:10; call sites main/backend/ba.py:35,39,44,49), which is not
installed here. Semantics: out = zeros(dim_size along `dim`); out.index_add_(dim, index, src).
"""
import torch


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    if dim < 0:
        dim = src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device) if out is None else out
    res.index_add_(dim, index.to(torch.long), src)
    return res
