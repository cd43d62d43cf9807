"""batrack_b200 — B200-native bundle-adjustment backend for BA-Track's sparse-SLAM hot path.

Drop-in surface (same names, arguments and return types as the reference):
    batrack_b200.ba.BA_rgbd_droid / BA         <- main/backend/ba.py:217 / :103
    batrack_b200.projective_ops.transform ...  <- main/backend/projective_ops.py:54
    batrack_b200.lietorch.SE3                  <- main/backend/lietorch/groups.py:266
    batrack_b200.lietorch_backends             <- the SE3 slice of the pybind extension (lietorch.cpp:286-316)
All of them run hand-written sm_100a CUDA through the C ABI in include/batrack_ba.h
(libbatrack_ba.so, built in-tree by __graft_entry__.build()). There is no CPU or eager fallback: a
missing library or a non-CUDA tensor raises.
"""
from ._capi import load_library, lib, launch_count  # noqa: F401

__all__ = ["load_library", "lib", "launch_count"]
