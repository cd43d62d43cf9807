"""Topology plans: the device-side structure that depends only on (ii, jj, kk), built once per graph
and reused by every BA call on that graph (the reference redoes this work inside every call:
main/backend/ba.py:219 `.item()` syncs, :269-277 index shift + torch.unique)."""
import ctypes as C
from collections import OrderedDict

import torch

from . import _capi


class Plan:
    """Owns a BaPlan* (include/batrack_ba.h). Keeps ii/jj/kk alive so a cache key can never be
    reused by different tensors at the same address."""

    def __init__(self, ii, jj, kk, n_poses, n_patches):
        for name, t in (("ii", ii), ("jj", jj), ("kk", kk)):
            if not isinstance(t, torch.Tensor) or not t.is_cuda:
                raise RuntimeError(f"{name}: batrack_b200 needs CUDA index tensors (no CPU fallback)")
            if t.dtype != torch.int64:
                raise TypeError(f"{name}: int64 required (got {t.dtype})")
        if not (ii.shape == jj.shape == kk.shape) or ii.dim() != 1:
            raise ValueError("ii, jj, kk must be 1-D tensors of equal length")
        self.device = ii.device
        self._keep = tuple(t if t.is_contiguous() else t.contiguous() for t in (ii, jj, kk))
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = _capi.lib().ba_plan_create(*[_capi.ptr(t) for t in self._keep], ii.numel(), int(n_poses),
                                            int(n_patches), _capi.stream_ptr(self.device), C.byref(handle))
        _capi.check(rc, "ba_plan_create")
        self.handle = handle
        self.info = _capi.BaPlanInfo()
        _capi.check(_capi.lib().ba_plan_info(self.handle, C.byref(self.info)), "ba_plan_info")
        self.layout_n_total = self.info.n_total

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _capi._lib is not None:
            _capi._lib.ba_plan_destroy(h)

    # ---- thin wrappers over the C ABI ----
    def tracks(self):
        """kx: sorted unique patch indices (== torch.unique(kk)), int32 [m]."""
        out = torch.empty(self.info.n_tracks, dtype=torch.int32, device=self.device)
        _capi.check(_capi.lib().ba_plan_tracks(self.handle, _capi.ptr(out), _capi.stream_ptr(self.device)))
        return out

    def set_option(self, name, value):
        """Per-plan switch (BA_OPT_* in include/batrack_ba.h): solver, stream, stream_smem_kb, schur_tile, twist_min,
        spin_cap, solver_trace, schur. `solver` also takes "auto" / "diag" / "tiles" / "mma" / "window" / "dense"."""
        if name == "solver" and isinstance(value, str):
            value = _capi.SOLVERS[value]
        _capi.check(_capi.lib().ba_plan_set_option(self.handle, _capi.OPTIONS[name], int(value)), f"set_option({name})")

    def get_option(self, name):
        v = C.c_int32()
        _capi.check(_capi.lib().ba_plan_get_option(self.handle, _capi.OPTIONS[name], C.byref(v)), f"get_option({name})")
        return int(v.value)

    def read_trace(self):
        """int64 clock stamps of the last traced band solve: [4096 columns][16 slots] + 4 phase stamps per side."""
        import numpy as np
        out = np.zeros(16 * 4096 + 32, dtype=np.int64)
        _capi.check(_capi.lib().ba_plan_read_trace(self.handle, out.ctypes.data_as(C.c_void_p), out.size,
                                                   _capi.stream_ptr(self.device)), "read_trace")
        return out

    def set_layout(self, n_total, block_bandwidth):
        _capi.check(_capi.lib().ba_plan_set_layout(self.handle, int(n_total), int(block_bandwidth)), "set_layout")
        self.layout_n_total = int(n_total)
        _capi.check(_capi.lib().ba_plan_info(self.handle, C.byref(self.info)))

    def reduced_system(self):
        """Zero-copy float64 tensor view of the exchange buffer [S | y] of the last assemble."""
        p, n = C.c_void_p(), C.c_int64()
        _capi.check(_capi.lib().ba_plan_reduced_system(self.handle, C.byref(p), C.byref(n)), "reduced_system")
        return _tensor_view(p.value, n.value, self.device, self, dtype=torch.float64)

    def debug(self, n):
        """Dense symmetric S [6n,6n], y, dX, Q, w, dZ of the last call (tests)."""
        m, M = self.info.n_tracks, 6 * n
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        out = dict(S=f(M, M), y=f(M), dX=f(M), Q=f(m), w=f(m), dZ=f(m))
        _capi.check(_capi.lib().ba_plan_debug_dense(
            self.handle, n, *[_capi.ptr(out[k]) if out[k].numel() else None for k in ("S", "y", "dX", "Q", "w", "dZ")],
            _capi.stream_ptr(self.device)), "debug_dense")
        return out

    def enable_timing(self, on=True):
        _capi.check(_capi.lib().ba_plan_enable_timing(self.handle, int(on)), "enable_timing")

    def last_timing(self):
        """{stage: milliseconds} of the last step (waits for it)."""
        buf = (C.c_float * len(_capi.STAGES))()
        _capi.check(_capi.lib().ba_plan_last_timing(self.handle, buf), "last_timing")
        return dict(zip(_capi.STAGES, [float(v) for v in buf]))

    def status(self):
        p = C.c_void_p()
        _capi.check(_capi.lib().ba_plan_status_ptr(self.handle, C.byref(p)))
        return int(_tensor_view(p.value, 1, self.device, self, dtype=torch.int32).item())


class CapacityPlan(Plan):
    """A plan allocated once for the largest graph a SLAM session will build. `update(ii, jj, kk)`
    re-derives it for the current graph on the device — no host synchronisation, no allocation; the counts come back
    through pinned memory and are read by the first BA call that uses the plan (`finalize()` does it explicitly).

    cap_groups: tracks with identical (ii, jj) lists form a group — one per source keyframe in BA-Track's graphs
    (main/batrack.py:399-410), so the number of frames that can source edges bounds it; cap_pattern bounds the sum over
    groups of the edges of one track. A graph that does not fit raises RuntimeError("...capacity...") at first use."""

    def __init__(self, n_poses, n_patches, cap_edges, cap_tracks=None, cap_groups=None, cap_pattern=None, cap_est=0,
                 device="cuda"):
        self.device = torch.device(device if isinstance(device, str) and ":" in device else torch.device(device, torch.cuda.current_device())
                                   if str(device) == "cuda" else device)
        cap_tracks = int(cap_tracks if cap_tracks is not None else min(cap_edges, n_patches))
        cap_groups = int(cap_groups if cap_groups is not None else min(cap_tracks, 4 * n_poses))
        cap_pattern = int(cap_pattern if cap_pattern is not None else min(cap_edges, 256 * cap_groups))
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = _capi.lib().ba_plan_create_capacity(int(cap_edges), cap_tracks, cap_groups, cap_pattern, int(cap_est),
                                                     int(n_poses), int(n_patches), _capi.stream_ptr(self.device), C.byref(handle))
        _capi.check(rc, "ba_plan_create_capacity")
        self.handle = handle
        self.info = _capi.BaPlanInfo()
        self._keep = ()
        self.layout_n_total = 0

    def update(self, ii, jj, kk, n_edges_dev=None):
        """Enqueue the re-derivation for (ii, jj, kk) on the current stream. n_edges_dev: optional int32 CUDA tensor with
        the live edge count (<= len(ii)) when the graph is maintained on the device."""
        for name, t in (("ii", ii), ("jj", jj), ("kk", kk)):
            if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.int64 or not t.is_contiguous():
                raise RuntimeError(f"{name}: contiguous int64 CUDA tensor required")
        self._keep = (ii, jj, kk, n_edges_dev)
        with torch.cuda.device(self.device):
            rc = _capi.lib().ba_plan_update(self.handle, _capi.ptr(ii), _capi.ptr(jj), _capi.ptr(kk), ii.numel(),
                                            _capi.ptr(n_edges_dev) if n_edges_dev is not None else None,
                                            _capi.stream_ptr(self.device))
        _capi.check(rc, "ba_plan_update")
        self._stale = True
        return self

    def finalize(self):
        _capi.check(_capi.lib().ba_plan_finalize(self.handle), "ba_plan_finalize")
        _capi.check(_capi.lib().ba_plan_info(self.handle, C.byref(self.info)), "ba_plan_info")
        self.layout_n_total = self.info.n_total
        self._stale = False
        return self


class _RawCuda:
    """__cuda_array_interface__ carrier for memory owned by the C library."""

    def __init__(self, ptr, n, typestr, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _tensor_view(ptr, n, device, owner, dtype=torch.float32):
    typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_RawCuda(ptr, n, typestr, owner), device=device)


_CACHE = OrderedDict()
_CACHE_MAX = 8


def get_plan(ii, jj, kk, n_poses, n_patches):
    """Plan for this topology, cached on (storage address, length, in-place version) of the three index
    tensors — main/batrack.py rebuilds them by torch.cat / boolean indexing when edges are added or
    removed (:196-198, :207-209) and edits them in place in keyframe() (:1049-1051); both change the key."""
    key = (ii.data_ptr(), jj.data_ptr(), kk.data_ptr(), ii.numel(), ii._version, jj._version, kk._version,
           int(n_poses), int(n_patches), str(ii.device))
    plan = _CACHE.get(key)
    if plan is None:
        plan = Plan(ii, jj, kk, n_poses, n_patches)
        plan._key_tensors = (ii, jj, kk)
        _CACHE[key] = plan
        while len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(key)
    return plan


def clear_plan_cache():
    _CACHE.clear()
