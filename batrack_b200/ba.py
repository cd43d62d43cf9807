"""Drop-in `BA_rgbd_droid` / `BA` (reference: main/backend/ba.py:217 / :103; call site
main/batrack.py:871-875). Same arguments, same return types (SE3 poses, fresh patches tensor), inputs
never modified. The whole damped Gauss-Newton step runs in libbatrack_ba.so (hand-written sm_100a
kernels, include/batrack_ba.h) on the current CUDA stream with no host synchronisation once the
topology plan of (ii, jj, kk) is cached.

Sharded graphs: pass `group=` (a torch.distributed process group over ranks that each
hold the edges of their own keyframe window, with replicated poses / intrinsics / patches): the only
exchange is one all-reduce of the reduced camera system [S | y].
"""
import ctypes as C

import torch

from . import _capi
from .lietorch import SE3
from .plan import get_plan


def _problem(plan, poses, patches, monodisp, intrinsics, targets, weights, lmbda, bounds, ep, fixedp,
             structure_only, loss, alpha):
    if loss not in _capi.LOSS_IDS:
        raise NotImplementedError(loss)                                   # ba.py:98-99
    if getattr(plan, "_stale", False):
        plan.finalize()                                                   # CapacityPlan.update(): read its counts now
    pdata = poses.data
    b, N = pdata.shape[0], pdata.shape[1]
    if b != 1:
        raise ValueError("batch size must be 1 (ba.py:218)")
    if patches.dim() != 5 or patches.shape[2] != 3 or patches.shape[3] != 1 or patches.shape[4] != 1:
        raise ValueError(f"patches must be [1, NM, 3, 1, 1] (P = 1, main/batrack.py:45); got {tuple(patches.shape)}")
    NM, E = patches.shape[1], plan.info.n_edges
    keep = []          # tensors that must outlive the launches

    def f32(name, t, shape):
        t = _capi.require_cuda_f32(name, t, contiguous=False)
        if tuple(t.shape) != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
        if not t.is_contiguous():
            t = t.contiguous()
        keep.append(t)
        return t

    p = _capi.BaProblem()
    p.poses = f32("poses", pdata, (1, N, 7)).data_ptr()
    p.patches = f32("patches", patches, (1, NM, 3, 1, 1)).data_ptr()
    p.monodisp = f32("patches_monodisp", monodisp, (1, NM, 1)).data_ptr() if monodisp is not None else None
    p.intrinsics = f32("intrinsics", intrinsics, (1, N, 4)).data_ptr()
    # targets: accept the strided view targets_3d[..., :2] the caller passes (main/batrack.py:871) as is
    tg = _capi.require_cuda_f32("targets", targets, contiguous=False)
    if tuple(tg.shape) != (1, E, 2):
        raise ValueError(f"targets: expected shape {(1, E, 2)}, got {tuple(tg.shape)}")
    if tg.stride(2) == 1 and tg.stride(1) in (2, 3) or E == 0:
        p.targets_stride = tg.stride(1)
    else:
        tg = tg.contiguous()
        p.targets_stride = 2
    keep.append(tg)
    p.targets = tg.data_ptr()
    p.weights = f32("weights", weights, (1, E, 2)).data_ptr()
    if isinstance(lmbda, torch.Tensor):
        m = plan.info.n_tracks
        if lmbda.numel() == 1:
            lmbda = lmbda.reshape(1).expand(m)                            # broadcasts like C + lmbda
        lv = _capi.require_cuda_f32("lmbda", lmbda.reshape(m), contiguous=False).contiguous()   # ba.py:299-300
        keep.append(lv)
        p.lmbda_vec = lv.data_ptr()
        p.lmbda = 0.0
    else:
        p.lmbda_vec = None
        p.lmbda = float(lmbda)
    p.ep, p.alpha = float(ep), float(alpha)
    for k in range(4):
        p.bounds[k] = float(bounds[k])
    p.fixedp, p.structure_only, p.loss = int(fixedp), int(bool(structure_only)), _capi.LOSS_IDS[loss]
    # a structure-only call returns the caller's poses object, like the reference (ba.py:336-339): nothing to allocate or copy
    poses_out = None if structure_only else torch.empty((1, N, 7), dtype=torch.float32, device=pdata.device)
    patches_out = torch.empty((1, NM, 3, 1, 1), dtype=torch.float32, device=pdata.device)
    p.poses_out, p.patches_out = (poses_out.data_ptr() if poses_out is not None else None), patches_out.data_ptr()
    return p, poses_out, patches_out, keep


def ensure_sharded_layout(plan, group):
    """Sharded calls sum the ranks' reduced systems element by element, so every rank must lay [S | y] out
    for the same number of poses and the same band width, and every rank must take part in the all-reduce: agree on
    max(n_total), max(block_bandwidth) once per (plan, group). Tracks must be partitioned by rank (each patch's edges on
    ONE rank): Q = 1 / (C + lambda) is formed from the local C."""
    if getattr(plan, "_layout_group", None) is group:
        return
    import torch.distributed as dist
    lay = torch.tensor([plan.info.n_total, plan.info.block_bandwidth], device=plan.device, dtype=torch.int64)
    dist.all_reduce(lay, op=dist.ReduceOp.MAX, group=group)
    n_total, bwb = int(lay[0]), int(lay[1])
    if n_total != plan.layout_n_total or bwb != plan.info.block_bandwidth:
        plan.set_layout(n_total, bwb)
    plan._layout_group = group


def _run(poses, patches, monodisp, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep, PRINT, fixedp,
         structure_only, loss, alpha, group=None, plan=None):
    pdata = poses.data
    if not pdata.is_cuda:
        raise RuntimeError("batrack_b200 runs on CUDA tensors only; there is no CPU fallback")
    if plan is None:
        plan = get_plan(ii, jj, kk, pdata.shape[1], patches.shape[1])
    prob, poses_out, patches_out, keep = _problem(plan, poses, patches, monodisp, intrinsics, targets, weights,
                                                  lmbda, bounds, ep, fixedp, structure_only, loss, alpha)
    if PRINT:                                                            # ba.py:244-245
        from . import projective_ops as pops
        coords, v = pops.transform(poses, patches, intrinsics, ii, jj, kk, valid=True)
        c, v = coords[..., 0, 0, :], v[..., 0, 0]                         # [1,E,2], [1,E] like ba.py:223-226
        r = targets - c
        v = v * (r.norm(dim=-1) < 250).float() * ((c[..., 0] > bounds[0]) & (c[..., 1] > bounds[1]) &
                                                  (c[..., 0] < bounds[2]) & (c[..., 1] < bounds[3])).float()
        print((r * v[..., None]).norm(dim=-1).mean().item())
    L = _capi.lib()
    with torch.cuda.device(pdata.device):
        st = _capi.stream_ptr(pdata.device)
        if group is None:
            _capi.check(L.ba_step(plan.handle, C.byref(prob), st), "ba_step")
        else:
            import torch.distributed as dist
            ensure_sharded_layout(plan, group)
            _capi.check(L.ba_assemble(plan.handle, C.byref(prob), st), "ba_assemble")
            if not structure_only and plan.layout_n_total - int(fixedp) > 0:
                dist.all_reduce(plan.reduced_system(), op=dist.ReduceOp.SUM, group=group)
            _capi.check(L.ba_solve_update(plan.handle, C.byref(prob), st), "ba_solve_update")
    del keep
    return (SE3(poses_out) if poses_out is not None else poses), patches_out


def BA_rgbd_droid(poses, patches, patches_monodisp, intrinsics, targets_2d, targets_disp, weights, lmbda, ii, jj, kk,
                  bounds, ep=100.0, PRINT=False, fixedp=1, structure_only=False, loss='trivial', alpha=0.5,
                  group=None, plan=None):
    """One damped Gauss-Newton step with the mono-disparity prior (ba.py:217-339). `targets_disp` is
    accepted and unused, as in the reference."""
    return _run(poses, patches, patches_monodisp, intrinsics, targets_2d, weights, lmbda, ii, jj, kk, bounds, ep,
                PRINT, fixedp, structure_only, loss, alpha, group, plan)


def BA(poses, patches, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep=100.0, PRINT=False, fixedp=1,
       structure_only=False, loss='trivial', group=None, plan=None):
    """One damped Gauss-Newton step without the prior (ba.py:103-213)."""
    return _run(poses, patches, None, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep, PRINT, fixedp,
                structure_only, loss, 0.0, group, plan)


def BA_update(poses, patches, patches_monodisp, intrinsics, targets_2d, weights_pose, weights, lmbda, ii, jj, kk, bounds,
              ep=10.0, fixedp=1, loss='huber', alpha=0.05, iters=4, plan=None):
    """The BA driver loop of BATRACK.update (main/batrack.py:869-875) as ONE native call:
    `iters` x { BA_rgbd_droid(weights_pose, structure_only=False); BA_rgbd_droid(weights, structure_only=True) },
    sharing the topology plan, the workspace and the launch stream, with no Python between the 2*iters steps.
    Equivalent to calling BA_rgbd_droid in that order; returns the final (SE3 poses, patches)."""
    pdata = poses.data
    if not pdata.is_cuda:
        raise RuntimeError("batrack_b200 runs on CUDA tensors only; there is no CPU fallback")
    if plan is None:
        plan = get_plan(ii, jj, kk, pdata.shape[1], patches.shape[1])
    prob, poses_out, patches_out, keep = _problem(plan, poses, patches, patches_monodisp, intrinsics, targets_2d,
                                                  weights_pose, lmbda, bounds, ep, fixedp, False, loss, alpha)
    E = plan.info.n_edges
    w_all = _capi.require_cuda_f32("weights", weights, contiguous=False)
    if tuple(w_all.shape) != (1, E, 2):
        raise ValueError(f"weights: expected shape {(1, E, 2)}, got {tuple(w_all.shape)}")
    w_all = w_all.contiguous()
    with torch.cuda.device(pdata.device):
        _capi.check(_capi.lib().ba_update(plan.handle, C.byref(prob), _capi.ptr(w_all), int(iters),
                                          _capi.stream_ptr(pdata.device)), "ba_update")
    del keep
    return (SE3(poses_out) if poses_out is not None else poses), patches_out
