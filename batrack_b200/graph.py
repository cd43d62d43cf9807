"""`FactorGraph`: BA-Track's edge bookkeeping (main/batrack.py:189-212 append_factors / remove_factors, :1023-1073 keyframe
edge surgery) on the device, behind the reference's method names. The edge list and its payload live in fixed capacity
buffers owned by libbatrack_ba.so; no operation synchronises with the host, and `plan()` re-derives the BA topology plan
for the current graph without synchronising either (CapacityPlan.update on buffers that never move: one graph launch)."""
import ctypes as C

import torch

from . import _capi
from .plan import CapacityPlan, _tensor_view, _RawCuda


def _view(ptr, shape, dtype, device, owner):
    n = 1
    for s in shape:
        n *= s
    typestr = {torch.int64: "<i8", torch.float32: "<f4", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_RawCuda(ptr, n, typestr, owner), device=device).view(*shape)


class FactorGraph:
    def __init__(self, n_poses, patches_per_frame, cap_edges, cap_groups=None, cap_pattern=None, device="cuda:0"):
        self.device = torch.device(device)
        self.N, self.M = int(n_poses), int(patches_per_frame)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().ba_graph_create(int(cap_edges), _capi.stream_ptr(self.device), C.byref(h)), "ba_graph_create")
        self.handle = h
        p = [C.c_void_p() for _ in range(7)]
        up, cap = C.c_int64(), C.c_int64()
        _capi.check(_capi.lib().ba_graph_arrays(self.handle, *[C.byref(x) for x in p], C.byref(up), C.byref(cap)))
        self.capacity = int(cap.value)
        c = self.capacity
        self._ii, self._jj, self._kk = (_view(x.value, (c,), torch.int64, self.device, self) for x in p[:3])
        self._tgt = _view(p[3].value, (c, 3), torch.float32, self.device, self)
        self._w = _view(p[4].value, (c, 2), torch.float32, self.device, self)
        self._wp = _view(p[5].value, (c, 2), torch.float32, self.device, self)
        self.n_edges_dev = _view(p[6].value, (1,), torch.int32, self.device, self)
        self._plan = CapacityPlan(self.N, self.N * self.M, cap_edges=c, cap_groups=cap_groups or self.N,
                                  cap_pattern=cap_pattern or 128 * (cap_groups or self.N), device=self.device)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _capi._lib is not None:
            self._plan = None
            _capi._lib.ba_graph_destroy(h)

    # ---- the reference's operations ----
    def append_factors(self, ii, jj, ix, targets_3d=None, weights=None, weights_pose=None):
        """main/batrack.py:189-204: ii = patch indices, jj = frame indices (int64 CUDA), ix = patch -> source frame table."""
        n = ii.numel()
        f = lambda t, k: None if t is None else _capi.ptr(_capi.require_cuda_f32("payload", t.reshape(n, k), contiguous=False).contiguous())
        keep = [ii.contiguous(), jj.contiguous(), ix.contiguous()]
        with torch.cuda.device(self.device):
            rc = _capi.lib().ba_graph_append(self.handle, _capi.ptr(keep[0]), _capi.ptr(keep[1]), n, _capi.ptr(keep[2]),
                                             f(targets_3d, 3), f(weights, 2), f(weights_pose, 2), _capi.stream_ptr(self.device))
        _capi.check(rc, "ba_graph_append")

    def remove_factors(self, mask):
        """main/batrack.py:206-212: mask [n_upper] bool / uint8 CUDA tensor, True removes."""
        m = mask.to(torch.uint8).contiguous()
        if m.numel() < self.n_upper:
            raise ValueError("mask shorter than the edge list")
        self._remove(0, 0, m, None)

    def remove_before(self, first_kept_frame, ix):
        """Removal window (main/batrack.py:1023-1026, :1072-1073): drop edges whose source frame ix[kk] < first_kept_frame."""
        self._remove(1, int(first_kept_frame), None, ix.contiguous())

    def remove_keyframe(self, k):
        """Edge part of keyframe() (main/batrack.py:1042-1051): drop the edges of frame k, shift the indices behind it."""
        self._remove(2, int(k), None, None)

    def _remove(self, mode, a, mask, ix):
        with torch.cuda.device(self.device):
            rc = _capi.lib().ba_graph_remove(self.handle, mode, a, self.M, _capi.ptr(mask) if mask is not None else None,
                                             _capi.ptr(ix) if ix is not None else None, _capi.stream_ptr(self.device))
        _capi.check(rc, "ba_graph_remove")

    # ---- views ----
    @property
    def n_upper(self):
        up = C.c_int64()
        _capi.check(_capi.lib().ba_graph_arrays(self.handle, None, None, None, None, None, None, None, C.byref(up), None))
        return int(up.value)

    def count(self):
        """Live edge count (synchronises the current stream)."""
        n = C.c_int64()
        _capi.check(_capi.lib().ba_graph_count(self.handle, C.byref(n), _capi.stream_ptr(self.device)), "ba_graph_count")
        return int(n.value)

    def plan(self):
        """Topology plan of the current graph: enqueued now, finalized (its counts read) at first use."""
        # always over the whole (fixed) buffers, the live count on the device: the same pointers and length every time, so
        # the derivation replays one captured graph
        return self._plan.update(self._ii, self._jj, self._kk, n_edges_dev=self.n_edges_dev)

    def edges(self):
        """(ii, jj, kk, targets_3d, weights, weights_pose) views of the live edges. Needs the count: taken from the plan's
        shape block when a plan update is in flight or done, else synchronises."""
        if getattr(self._plan, "_stale", False):
            self._plan.finalize()
            n = self._plan.info.n_edges
            _capi.check(_capi.lib().ba_graph_tighten(self.handle, n))
        else:
            n = self.count()
        return self._ii[:n], self._jj[:n], self._kk[:n], self._tgt[:n], self._w[:n], self._wp[:n]
