"""Drop-in slice of main/backend/projective_ops.py: `transform` without Jacobians (the form the BA
caller's neighbours use: reproject main/batrack.py:327-338, flow_mag projective_ops.py:112-122), plus
the cheap elementwise helpers `iproj` / `proj`. The Jacobian form lives fused inside the BA edge pass
(csrc/ba_kernels.cu) and is never materialised."""
import torch

from . import _capi

MIN_DEPTH = 0.2


def iproj(patches, intrinsics):
    """projective_ops.py:19-29"""
    x, y, d = patches.unbind(dim=2)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    return torch.stack([(x - cx) / fx, (y - cy) / fy, torch.ones_like(d), d], dim=-1)


def proj(X, intrinsics, depth=False):
    """projective_ops.py:32-52"""
    X, Y, Z, W = X.unbind(dim=-1)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    d = 1.0 / Z.clamp(min=1e-2)
    x, y = fx * (d * X) + cx, fy * (d * Y) + cy
    return torch.stack([x, y, d * W], dim=-1) if depth else torch.stack([x, y], dim=-1)


def transform(poses, patches, intrinsics, ii, jj, kk, depth=False, valid=False, jacobian=False, tonly=False):
    """projective_ops.py:54-105 for P = 1 patches. Returns coords [1,E,1,1,2] (and valid [1,E,1,1] when
    `valid`). `jacobian=True` / `depth=True` are not exposed: the fused BA kernel owns that path."""
    if jacobian or depth:
        raise NotImplementedError("transform(jacobian/depth=True) is fused into BA_rgbd_droid in batrack_b200")
    pdata = _capi.require_cuda_f32("poses", poses.data, contiguous=False).contiguous()
    pt = _capi.require_cuda_f32("patches", patches, contiguous=False).contiguous()
    K = _capi.require_cuda_f32("intrinsics", intrinsics, contiguous=False).contiguous()
    if pt.dim() != 5 or pt.shape[3] != 1 or pt.shape[4] != 1:
        raise ValueError("patches must be [1, NM, 3, 1, 1]")
    idx = [t.contiguous() for t in (ii, jj, kk)]
    for t in idx:
        if not t.is_cuda or t.dtype != torch.int64:
            raise TypeError("ii, jj, kk must be int64 CUDA tensors")
    E = idx[0].numel()
    coords = torch.empty((1, E, 1, 1, 2), dtype=torch.float32, device=pdata.device)
    v = torch.empty((1, E, 1, 1), dtype=torch.float32, device=pdata.device) if valid else None
    with torch.cuda.device(pdata.device):
        rc = _capi.lib().ba_reproject(_capi.ptr(pdata), _capi.ptr(pt), _capi.ptr(K), *[_capi.ptr(t) for t in idx], E,
                                      pdata.shape[1], pt.shape[1], int(bool(tonly)), _capi.ptr(coords), _capi.ptr(v),
                                      _capi.stream_ptr(pdata.device))
    _capi.check(rc, "ba_reproject")
    return (coords, v) if valid else coords


def flow_mag(poses, patches, intrinsics, ii, jj, kk, beta=0.3):
    """projective_ops.py:112-122"""
    c0 = transform(poses, patches, intrinsics, ii, ii, kk)
    c1 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=False)
    c2 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=True)
    return beta * (c1 - c0).norm(dim=-1) + (1 - beta) * (c2 - c0).norm(dim=-1)


def point_cloud(poses, patches, intrinsics, ix):
    """projective_ops.py:107-109: back-projected patches in the world frame, [1, M, P, P, 4] homogeneous points
    (x, y, z, inverse depth). SE3 inverse / action run in the library's SE3 kernels."""
    return poses[:, ix, None, None].inv() * iproj(patches, intrinsics[:, ix])
