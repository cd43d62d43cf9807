"""Drop-in `main/backend/projective_ops.py`: `transform` (all of its forms: plain, `valid`, `depth`, `jacobian`, `tonly`),
`iproj`, `proj`, `point_cloud`, `flow_mag`, `back_proj`, `proj_to_frames` with the reference's signatures and shapes, for
P = 1 patches (main/batrack.py:45). Everything that touches poses runs in libbatrack_ba.so (csrc/ba_aux.cu); the BA
iteration itself never calls these — its Jacobians live fused inside the edge pass and are not materialised."""
import torch

from . import _capi

MIN_DEPTH = 0.2


def extract_intrinsics(intrinsics):
    """projective_ops.py:9-10"""
    return intrinsics[..., None, None, :].unbind(dim=-1)


def coords_grid(ht, wd, **kwargs):
    """projective_ops.py:12-17"""
    y, x = torch.meshgrid(torch.arange(ht).to(**kwargs).float(), torch.arange(wd).to(**kwargs).float(), indexing="ij")
    return torch.stack([x, y], dim=-1)


def iproj(patches, intrinsics):
    """projective_ops.py:19-29 (elementwise; no pose involved)"""
    x, y, d = patches.unbind(dim=2)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    return torch.stack([(x - cx) / fx, (y - cy) / fy, torch.ones_like(d), d], dim=-1)


def proj(X, intrinsics, depth=False):
    """projective_ops.py:32-52 (elementwise; no pose involved)"""
    X, Y, Z, W = X.unbind(dim=-1)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    d = 1.0 / Z.clamp(min=1e-2)
    x, y = fx * (d * X) + cx, fy * (d * Y) + cy
    return torch.stack([x, y, d * W], dim=-1) if depth else torch.stack([x, y], dim=-1)


def _inputs(poses, patches, intrinsics, ii, jj, kk):
    pdata = _capi.require_cuda_f32("poses", poses.data, contiguous=False).contiguous()
    pt = _capi.require_cuda_f32("patches", patches, contiguous=False).contiguous()
    K = _capi.require_cuda_f32("intrinsics", intrinsics, contiguous=False).contiguous()
    if pt.dim() != 5 or pt.shape[3] != 1 or pt.shape[4] != 1:
        raise ValueError("patches must be [1, NM, 3, 1, 1]")
    idx = [t.contiguous() for t in (ii, jj, kk)]
    for t in idx:
        if not t.is_cuda or t.dtype != torch.int64:
            raise TypeError("ii, jj, kk must be int64 CUDA tensors")
    return pdata, pt, K, idx


def transform(poses, patches, intrinsics, ii, jj, kk, depth=False, valid=False, jacobian=False, tonly=False):
    """projective_ops.py:54-105 for P = 1 patches. Returns coords [1,E,1,1,2] ([...,3] with `depth`);
    with `valid`: (coords, valid [1,E,1,1]); with `jacobian`: (coords, valid [1,E], (Ji [1,E,2,6], Jj [1,E,2,6],
    Jz [1,E,2,1])) exactly like the reference (which returns the Z > 0.2 mask in that form whatever `valid` says)."""
    pdata, pt, K, idx = _inputs(poses, patches, intrinsics, ii, jj, kk)
    E, dev = idx[0].numel(), pdata.device
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    coords = f(1, E, 1, 1, 3 if depth else 2)
    v = f(1, E) if (valid or jacobian) else None
    Ji, Jj, Jz = (f(1, E, 2, 6), f(1, E, 2, 6), f(1, E, 2, 1)) if jacobian else (None, None, None)
    with torch.cuda.device(dev):
        rc = _capi.lib().ba_transform(_capi.ptr(pdata), _capi.ptr(pt), _capi.ptr(K), *[_capi.ptr(t) for t in idx], E,
                                      pdata.shape[1], pt.shape[1], int(bool(tonly)), int(bool(depth)), _capi.ptr(coords),
                                      _capi.ptr(v), _capi.ptr(Ji), _capi.ptr(Jj), _capi.ptr(Jz), _capi.stream_ptr(dev))
    _capi.check(rc, "ba_transform")
    if jacobian:
        return coords, v, (Ji, Jj, Jz)
    if valid:
        return coords, v.view(1, E, 1, 1)
    return coords


def flow_mag(poses, patches, intrinsics, ii, jj, kk, beta=0.3):
    """projective_ops.py:112-122"""
    c0 = transform(poses, patches, intrinsics, ii, ii, kk)
    c1 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=False)
    c2 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=True)
    return beta * (c1 - c0).norm(dim=-1) + (1 - beta) * (c2 - c0).norm(dim=-1)


def point_cloud(poses, patches, intrinsics, ix):
    """projective_ops.py:107-109: back-projected patches in the world frame, [1, M, 1, 1, 4] homogeneous points
    (x, y, z, inverse depth); one fused kernel (pose inverse + inverse projection + action)."""
    pdata = _capi.require_cuda_f32("poses", poses.data, contiguous=False).contiguous()
    pt = _capi.require_cuda_f32("patches", patches, contiguous=False).contiguous()
    K = _capi.require_cuda_f32("intrinsics", intrinsics, contiguous=False).contiguous()
    if pt.dim() != 5 or pt.shape[3] != 1 or pt.shape[4] != 1:
        raise ValueError("patches must be [1, M, 3, 1, 1]")
    ix = ix.contiguous()
    if not ix.is_cuda or ix.dtype != torch.int64 or ix.numel() != pt.shape[1]:
        raise TypeError("ix must be an int64 CUDA tensor with one frame index per patch")
    n = pt.shape[1]
    out = torch.empty((1, n, 1, 1, 4), dtype=torch.float32, device=pdata.device)
    with torch.cuda.device(pdata.device):
        rc = _capi.lib().ba_point_cloud(_capi.ptr(pdata), _capi.ptr(pt), _capi.ptr(K), _capi.ptr(ix), n, pdata.shape[1],
                                        _capi.ptr(out), _capi.stream_ptr(pdata.device))
    _capi.check(rc, "ba_point_cloud")
    return out


def back_proj(xy, xy_depth, intrinsics, cams_c2w=None):
    """projective_ops.py:129-152: xy [B,N,2], xy_depth [B,N,1], intrinsics [B,4], cams_c2w [B,4,4] | None -> P [B,N,4]"""
    xy = _capi.require_cuda_f32("xy", xy, contiguous=False).contiguous()
    D = _capi.require_cuda_f32("xy_depth", xy_depth, contiguous=False).contiguous()
    K = _capi.require_cuda_f32("intrinsics", intrinsics, contiguous=False).contiguous()
    B, n = xy.shape[0], xy.shape[1]
    T = None if cams_c2w is None else cams_c2w.float().contiguous()
    if T is not None and not T.is_cuda:
        raise RuntimeError("cams_c2w: CUDA tensor required (no CPU fallback)")
    P = torch.empty((B, n, 4), dtype=torch.float32, device=xy.device)
    with torch.cuda.device(xy.device):
        rc = _capi.lib().ba_back_proj(_capi.ptr(xy), _capi.ptr(D), _capi.ptr(K), _capi.ptr(T), B, n, _capi.ptr(P),
                                      _capi.stream_ptr(xy.device))
    _capi.check(rc, "ba_back_proj")
    return P


def proj_to_frames(P, intrinsics, cams_w2c):
    """projective_ops.py:154-176: P [B,N,4], intrinsics [B,S,4], cams_w2c [B,S,4,4] -> xy [B,S,N,2]"""
    P = _capi.require_cuda_f32("P", P, contiguous=False).contiguous()
    K = _capi.require_cuda_f32("intrinsics", intrinsics, contiguous=False).contiguous()
    T = cams_w2c.float().contiguous()
    if not T.is_cuda:
        raise RuntimeError("cams_w2c: CUDA tensor required (no CPU fallback)")
    B, n, S = P.shape[0], P.shape[1], T.shape[1]
    xy = torch.empty((B, S, n, 2), dtype=torch.float32, device=P.device)
    with torch.cuda.device(P.device):
        rc = _capi.lib().ba_proj_to_frames(_capi.ptr(P), _capi.ptr(K), _capi.ptr(T), B, S, n, _capi.ptr(xy),
                                           _capi.stream_ptr(P.device))
    _capi.check(rc, "ba_proj_to_frames")
    return xy
