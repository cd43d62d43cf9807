"""Host-buffer front end of the BA step (include/batrack_ba.h: ba_step_host_async / ba_host_sync).

A caller that keeps its state in CPU arrays (pinned torch tensors) submits BA_rgbd_droid-shaped steps; the
library uploads the inputs of step k+1 and downloads the results of step k-1 while the kernels of step k run
(two device staging slots, its own copy streams). Arguments mean what they mean for `BA_rgbd_droid`
(main/backend/ba.py:217); `poses` is the raw [1,N,7] array, results land in caller-provided host tensors."""
import ctypes as C

import torch

from . import _capi


def _host_f32(name, t, shape):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if t.is_cuda:
        raise RuntimeError(f"{name}: HostBA takes CPU tensors (use batrack_b200.ba.BA_rgbd_droid for CUDA tensors)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: float32 required (got {t.dtype})")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: input must be contiguous")
    return t


class HostBA:
    """Pipelined BA steps on host arrays for one topology plan (batrack_b200.plan.Plan)."""

    def __init__(self, plan):
        self.plan = plan
        self._keep = []

    def submit(self, poses, patches, patches_monodisp, intrinsics, targets_2d, weights, lmbda, bounds, poses_out,
               patches_out, ep=100.0, fixedp=1, structure_only=False, loss='trivial', alpha=0.5, group=None):
        """Enqueue one step on the current CUDA stream and return at once. Inputs must stay unchanged and outputs
        are valid only after `sync()`. An input array that an earlier, still running step writes its results into
        (iteration k+1 starting from the results of iteration k, main/batrack.py:869-884) is uploaded after that
        download, ordered on the device: dependent steps can be submitted back to back without a `sync()` in between. With `group` (keyframe-sharded graph) the step is staged, assembled,
        all-reduced over the group and solved through ba_stage_host_async / ba_assemble / ba_solve_update /
        ba_unstage_host_async."""
        if loss not in _capi.LOSS_IDS:
            raise NotImplementedError(loss)
        if getattr(self.plan, "_stale", False):
            self.plan.finalize()                          # CapacityPlan.update(): read its counts now
        info = self.plan.info
        N, NM, E = info.n_poses, info.n_patches, info.n_edges
        p = _capi.BaProblem()
        p.poses = _host_f32("poses", poses, (1, N, 7)).data_ptr()
        p.patches = _host_f32("patches", patches, (1, NM, 3, 1, 1)).data_ptr()
        p.monodisp = _host_f32("patches_monodisp", patches_monodisp, (1, NM, 1)).data_ptr() \
            if patches_monodisp is not None else None
        p.intrinsics = _host_f32("intrinsics", intrinsics, (1, N, 4)).data_ptr()
        p.targets = _host_f32("targets_2d", targets_2d, (1, E, 2)).data_ptr()
        p.targets_stride = 2
        p.weights = _host_f32("weights", weights, (1, E, 2)).data_ptr()
        if isinstance(lmbda, torch.Tensor):
            lv = _host_f32("lmbda", lmbda.reshape(-1).expand(info.n_tracks).contiguous(), (info.n_tracks,))
            self._keep.append(lv)
            p.lmbda_vec, p.lmbda = lv.data_ptr(), 0.0
        else:
            p.lmbda_vec, p.lmbda = None, float(lmbda)
        p.ep, p.alpha = float(ep), float(alpha)
        for k in range(4):
            p.bounds[k] = float(bounds[k])
        p.fixedp, p.structure_only, p.loss = int(fixedp), int(bool(structure_only)), _capi.LOSS_IDS[loss]
        p.poses_out = _host_f32("poses_out", poses_out, (1, N, 7)).data_ptr()
        p.patches_out = _host_f32("patches_out", patches_out, (1, NM, 3, 1, 1)).data_ptr()
        dev = self.plan.device
        L = _capi.lib()
        with torch.cuda.device(dev):
            st = _capi.stream_ptr(dev)
            if group is None:
                _capi.check(L.ba_step_host_async(self.plan.handle, C.byref(p), st), "ba_step_host_async")
            else:
                import torch.distributed as dist
                from .ba import ensure_sharded_layout
                ensure_sharded_layout(self.plan, group)
                pd = _capi.BaProblem()
                _capi.check(L.ba_stage_host_async(self.plan.handle, C.byref(p), C.byref(pd), st), "ba_stage_host_async")
                _capi.check(L.ba_assemble(self.plan.handle, C.byref(pd), st), "ba_assemble")
                if not structure_only and self.plan.layout_n_total - int(fixedp) > 0:
                    dist.all_reduce(self.plan.reduced_system(), op=dist.ReduceOp.SUM, group=group)
                _capi.check(L.ba_solve_update(self.plan.handle, C.byref(pd), st), "ba_solve_update")
                _capi.check(L.ba_unstage_host_async(self.plan.handle, C.byref(p), C.byref(pd), st), "ba_unstage_host_async")

    def prefetch(self, patches_monodisp=None, intrinsics=None, targets_2d=None, weights=None, lmbda=None):
        """Upload the step-independent inputs of the NEXT submit() now (ba_prefetch_host_async): they travel while the
        previous step still computes, and the next submit() — which must pass the same arrays — uploads only poses and
        patches. For dependent steps: submit(k); prefetch(inputs of k+1); sync(); submit(k+1, poses / patches of k)."""
        info = self.plan.info
        N, NM, E = info.n_poses, info.n_patches, info.n_edges
        p = _capi.BaProblem()
        if patches_monodisp is not None:
            p.monodisp = _host_f32("patches_monodisp", patches_monodisp, (1, NM, 1)).data_ptr()
        if intrinsics is not None:
            p.intrinsics = _host_f32("intrinsics", intrinsics, (1, N, 4)).data_ptr()
        if targets_2d is not None:
            p.targets = _host_f32("targets_2d", targets_2d, (1, E, 2)).data_ptr()
            p.targets_stride = 2
        if weights is not None:
            p.weights = _host_f32("weights", weights, (1, E, 2)).data_ptr()
        if isinstance(lmbda, torch.Tensor):
            lv = _host_f32("lmbda", lmbda.reshape(-1).expand(info.n_tracks).contiguous(), (info.n_tracks,))
            self._keep.append(lv)
            p.lmbda_vec = lv.data_ptr()
        dev = self.plan.device
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().ba_prefetch_host_async(self.plan.handle, C.byref(p), _capi.stream_ptr(dev)), "ba_prefetch_host_async")

    def sync(self, block=True):
        """Order the current stream after every submitted step's download; block=True also waits on the host."""
        dev = self.plan.device
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().ba_host_sync(self.plan.handle, _capi.stream_ptr(dev), int(bool(block))),
                        "ba_host_sync")
        if block:
            self._keep.clear()
