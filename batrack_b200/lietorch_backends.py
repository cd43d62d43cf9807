"""SE3 slice of the reference's `lietorch_backends` pybind module, same call convention:
`fn(group_id, *contiguous 2-D float32 / float64 CUDA tensors) -> tensor`
(main/backend/lietorch/src/lietorch.cpp:286-316). Group ids as in lietorch/groups.py:236-290; only
SE3 (3) is on BA-Track's BA path, the other groups and all *_backward entry points are out of scope
 and raise. Every call launches a hand-written kernel from libbatrack_ba.so on
the current CUDA stream.
"""
import torch

from . import _capi

SE3_ID = 3


def _run(name, gid, out_cols, X, *others):
    if gid != SE3_ID:
        raise NotImplementedError(f"lietorch_backends.{name}: only SE3 (group id 3) is built, got {gid}")
    ts = (X,) + others
    for i, t in enumerate(ts):
        _capi.require_cuda_float(f"{name} arg{i}", t)
        if t.dtype != X.dtype:
            raise TypeError(f"{name}: mixed dtypes {[x.dtype for x in ts]}")
        if t.dim() != 2 or t.shape[0] != X.shape[0]:
            raise RuntimeError(f"{name}: expected 2-D tensors with equal batch, got {[tuple(x.shape) for x in ts]}")
    B = X.shape[0]
    out = torch.empty((B, out_cols), dtype=X.dtype, device=X.device)
    with torch.cuda.device(X.device):
        fn = getattr(_capi.lib(), ("se3d_" if X.dtype == torch.float64 else "se3_") + name)   # dispatch.h:37-45
        rc = fn(*[_capi.ptr(t) for t in ts], _capi.ptr(out), B, _capi.stream_ptr(X.device))
    _capi.check(rc, name)
    return out


def expm(gid, a):
    return _run("expm", gid, 7, a)


def logm(gid, X):
    return _run("logm", gid, 6, X)


def inv(gid, X):
    return _run("inv", gid, 7, X)


def mul(gid, X, Y):
    return _run("mul", gid, 7, X, Y)


def adj(gid, X, a):
    return _run("adj", gid, 6, X, a)


def adjT(gid, X, a):
    return _run("adjT", gid, 6, X, a)


def act(gid, X, p):
    return _run("act", gid, 3, X, p)


def act4(gid, X, p):
    return _run("act4", gid, 4, X, p)


def as_matrix(gid, X):
    return _run("as_matrix", gid, 16, X).view(-1, 4, 4)


def _no_backward(*_a, **_k):
    raise NotImplementedError("batrack_b200.lietorch_backends: backward ops are out of scope "
                              "(BA runs on detached tensors, main/batrack.py:871-875)")


expm_backward = logm_backward = inv_backward = mul_backward = _no_backward
adj_backward = adjT_backward = act_backward = act4_backward = _no_backward
projector = Jinv = _no_backward
