"""Trajectory / results hand-off: what BATRACK.terminate (main/batrack.py:898-915) and
BATRACK.get_results (:1080-1135) compute from the keyframe buffers — every processed frame's pose (dropped keyframes are
re-attached through their stored relative pose, get_pose :223-228), inverted to world-from-camera — in ONE kernel launch
(`ba_trajectory`) instead of `counter` Python-level SE3 products, plus the `results.pkl` dictionary global_refine reads.
"""
import pickle

import numpy as np
import torch

from . import _capi
from .lietorch import SE3


def _chain_tables(tstamps, n, delta, counter, device):
    """slot[t]: keyframe row holding frame t's pose or -1; t0[t] / dP[t]: the delta table (main/batrack.py:1037-1040)."""
    slot = np.full(counter, -1, dtype=np.int32)
    ts = tstamps[:n].detach().cpu().numpy().astype(np.int64) if isinstance(tstamps, torch.Tensor) else np.asarray(tstamps[:n], dtype=np.int64)
    for i, t in enumerate(ts):
        if 0 <= t < counter:
            slot[t] = i                                   # later rows win, like the dict assignment at :900-901
    t0 = np.full(counter, -1, dtype=np.int32)
    dP = np.zeros((counter, 7), dtype=np.float32)
    dP[:, 6] = 1.0
    for t, (src, d) in delta.items():
        t = int(t)
        if 0 <= t < counter and slot[t] < 0:              # `if t in self.traj` is checked first (:224-225)
            t0[t] = int(src)
            dP[t] = (d.data if isinstance(d, SE3) else d).detach().reshape(7).cpu().numpy()
    f = lambda a: torch.from_numpy(a).to(device)
    return f(slot), f(t0), f(dP)


def trajectory(poses, tstamps, n, delta, counter, want="both"):
    """poses [N,7] keyframe buffer (poses_), tstamps [N] (tstamps_), n live keyframes, delta {t: (t0, SE3 dP)}, counter
    frames processed. Returns (poses7 [counter,7] as [tx ty tz qw qx qy qz], cams_T_world [counter,4,4]) CUDA tensors."""
    p = _capi.require_cuda_f32("poses", poses.reshape(-1, 7), contiguous=False).contiguous()
    dev = p.device
    slot, t0, dP = _chain_tables(tstamps, n, delta, counter, dev)
    out7 = torch.empty(counter, 7, device=dev) if want in ("both", "vec") else None
    out44 = torch.empty(counter, 4, 4, device=dev) if want in ("both", "matrix") else None
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _capi.lib().ba_trajectory(_capi.ptr(p), _capi.ptr(slot), _capi.ptr(t0), _capi.ptr(dP), int(counter),
                                       _capi.ptr(out7) if out7 is not None else None,
                                       _capi.ptr(out44) if out44 is not None else None, _capi.ptr(err), _capi.stream_ptr(dev))
    _capi.check(rc, "ba_trajectory")
    if int(err.item()):
        raise KeyError("trajectory: a frame has neither a keyframe pose nor a delta (main/batrack.py:227 would raise KeyError)")
    return out7, out44


def terminate(poses, tstamps, n, delta, counter, tlist):
    """main/batrack.py:898-915: (poses [counter,7] numpy in [tx ty tz qw qx qy qz] order, tstamps float numpy)."""
    out7, _ = trajectory(poses, tstamps, n, delta, counter, want="vec")
    return out7.cpu().numpy(), np.array(tlist, dtype=float)


def get_results(poses, tstamps, n, delta, counter, tlist, intrinsics, patches_valid, patches_local, patches_local_weights,
                patches_local_static, patches_local_vis, rgbs=None, dmaps=None, dmaps_gt=None, save_path=None):
    """main/batrack.py:1080-1135: the dictionary global_refine consumes (same keys, shapes and dtypes), optionally pickled."""
    _, cams = trajectory(poses, tstamps, n, delta, counter, want="matrix")
    pts_valid = patches_valid[:counter].detach().cpu().numpy()
    trajs_valid = patches_local_weights[:counter, ..., 0]
    results = {
        "cams_T_world": cams.cpu().numpy(),
        "intrinsics": intrinsics[:counter].detach().cpu().numpy(),
        "tstamps": np.array(tlist, dtype=float),
        "trajs_2d_disp": patches_local[:counter].detach().cpu().numpy(),
        "trajs_valid": (trajs_valid.sum(dim=2) > 0).detach().cpu().numpy(),
        "trajs_static": patches_local_static[:counter, ..., 0].detach().cpu().numpy(),
        "trajs_vis": patches_local_vis[:counter, ..., 0].detach().cpu().numpy(),
        "grid_query_frames": np.arange(counter)[pts_valid.sum(axis=1) > 0],
        "dmaps": None if dmaps is None else np.array(dmaps, dtype=float),
        "rgbs": None if rgbs is None else np.array(rgbs, dtype=float),
        "dmaps_gt": None if dmaps_gt is None else np.array(dmaps_gt, dtype=float),
    }
    if save_path is not None:
        with open(save_path, "wb+") as f:
            pickle.dump(results, f)
    return results
