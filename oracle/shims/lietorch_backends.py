"""ORACLE SHIM (test infrastructure): stand-in for the reference's `lietorch_backends` extension.

The real extension (reference setup.py:17-24, main/backend/lietorch/src/lietorch.cpp:286-316)
needs Eigen 3.4.0, which is neither vendored nor installed, so it cannot be built here.
Putting this directory first on sys.path lets the reference's own Python BA code
(main/backend/ba.py, projective_ops.py, lietorch/groups.py, group_ops.py, broadcasting.py)
be imported UNMODIFIED on CPU; see tests/golden/make_golden.py.

Only SE3 (group id 3) forward ops are provided. `group_ops.py:28-66` merely binds the
`*_backward` attributes at import time, so those may be None.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import se3_ops as _s  # noqa: E402

_SE3 = 3


def _chk(gid, *ts):
    if gid != _SE3:
        raise NotImplementedError("oracle shim implements SE3 (group id 3) only")
    for t in ts:
        if not t.is_contiguous():
            raise RuntimeError("input must be contiguous")       # lietorch.cpp:7


def expm(gid, a):
    _chk(gid, a); return _s.se3_exp(a)


def logm(gid, X):
    _chk(gid, X); return _s.se3_log(X)


def inv(gid, X):
    _chk(gid, X); return _s.se3_inv(X)


def mul(gid, X, Y):
    _chk(gid, X, Y); return _s.se3_mul(X, Y)


def adj(gid, X, a):
    _chk(gid, X, a); return _s.se3_adj(X, a)


def adjT(gid, X, a):
    _chk(gid, X, a); return _s.se3_adjT(X, a)


def act(gid, X, p):
    _chk(gid, X, p); return _s.se3_act3(X, p)


def act4(gid, X, p):
    _chk(gid, X, p); return _s.se3_act4(X, p)


def as_matrix(gid, X):
    _chk(gid, X); return _s.se3_matrix(X)


expm_backward = logm_backward = inv_backward = mul_backward = None
adj_backward = adjT_backward = act_backward = act4_backward = None
projector = Jinv = None
