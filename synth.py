"""Seeded synthetic factor graphs for tests and bench.py (SURVEY.md §8(d) recipe).

Everything is generated on the CPU in float64 with numpy and cast to float32 / int64 at the end,
so the same call gives bit-identical inputs here and on the GPU box.

Graph conventions follow the caller of the BA operator (reference main/batrack.py:189-204,
399-410): `kk[e]` = patch (track) index, `ii[e]` = source frame of that patch, `jj[e]` = frame
the patch is reprojected into; edges are patch-major / frame-minor; every track also carries its
self-edge (jj == ii), exactly as `flatmeshgrid(kf_idx, frames)` produces.
"""
from dataclasses import dataclass, field

import numpy as np


# ---- minimal fp64 SE3 helpers (generator-private; [tx,ty,tz,qx,qy,qz,qw]) ---------------------

def _hat(p):
    z = np.zeros(p.shape[0])
    return np.stack([z, -p[:, 2], p[:, 1], p[:, 2], z, -p[:, 0], -p[:, 1], p[:, 0], z], 1).reshape(-1, 3, 3)


def _exp(xi):
    tau, phi = xi[:, :3], xi[:, 3:]
    th = np.linalg.norm(phi, axis=1)
    th2 = th * th
    small = th < 1e-6
    s = np.where(small, 1.0, th)
    imag = np.where(small, 0.5 - th2 / 48.0, np.sin(0.5 * s) / s)
    real = np.where(small, 1.0 - th2 / 8.0, np.cos(0.5 * s))
    q = np.concatenate([imag[:, None] * phi, real[:, None]], 1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    c1 = np.where(small, 0.5 - th2 / 24.0, (1 - np.cos(s)) / (s * s))
    c2 = np.where(small, 1 / 6.0 - th2 / 120.0, (s - np.sin(s)) / (s * s * s))
    P = _hat(phi)
    V = np.eye(3)[None] + c1[:, None, None] * P + c2[:, None, None] * (P @ P)
    t = (V @ tau[:, :, None])[:, :, 0]
    return np.concatenate([t, q], 1)


def _rot(q, p):
    qv, w = q[:, :3], q[:, 3:4]
    uv = 2.0 * np.cross(qv, p)
    return p + w * uv + np.cross(qv, uv)


def _qmul(a, b):
    ax, ay, az, aw = a.T
    bx, by, bz, bw = b.T
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], 1)


def _mul(X, Y):
    q = _qmul(X[:, 3:], Y[:, 3:])
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([X[:, :3] + _rot(X[:, 3:], Y[:, :3]), q], 1)


def _inv(X):
    qi = X[:, 3:] * np.array([-1.0, -1.0, -1.0, 1.0])
    return np.concatenate([-_rot(qi, X[:, :3]), qi], 1)


def _reproject(poses, xyd, intr, ii, jj, kk):
    """pixel of patch kk (seen in frame ii) in frame jj -> [E,2], plus depth Z in frame jj."""
    Gij = _mul(poses[jj], _inv(poses[ii]))
    fx, fy, cx, cy = intr[ii].T
    x, y, d = xyd[kk].T
    X0 = np.stack([(x - cx) / fx, (y - cy) / fy, np.ones_like(x)], 1)
    X1 = _rot(Gij[:, 3:], X0) + Gij[:, :3] * d[:, None]
    fx, fy, cx, cy = intr[jj].T
    z = np.maximum(X1[:, 2], 1e-2)
    return np.stack([fx * X1[:, 0] / z + cx, fy * X1[:, 1] / z + cy], 1), X1[:, 2]


# ---- problem container ------------------------------------------------------------------------

@dataclass
class BAProblem:
    """Host-side (numpy) inputs of one BA call, in the shapes main/backend/ba.py:217 expects
    once wrapped as torch tensors (see `as_torch`)."""
    poses: np.ndarray            # [N,7]  f32  initial estimate
    patches: np.ndarray          # [N*M,3] f32 (x, y, inverse depth); P = 1
    monodisp: np.ndarray         # [N*M]  f32
    intrinsics: np.ndarray       # [N,4]  f32 (fx, fy, cx, cy)
    targets: np.ndarray          # [E,2]  f32
    weights: np.ndarray          # [E,2]  f32
    ii: np.ndarray               # [E] i64
    jj: np.ndarray               # [E] i64
    kk: np.ndarray               # [E] i64
    bounds: list
    fixedp: int = 1
    ep: float = 10.0
    lmbda: float = 1e-4
    alpha: float = 0.05
    loss: str = "huber"
    name: str = ""
    gt_poses: np.ndarray = field(default=None, repr=False)
    gt_disp: np.ndarray = field(default=None, repr=False)

    @property
    def E(self):
        return int(self.ii.shape[0])

    def as_torch(self, device="cpu", dtype=None):
        """-> dict of tensors shaped like the reference's call site (main/batrack.py:864-875)."""
        import torch
        dt = dtype or torch.float32
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
        g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device=device)
        NM = self.patches.shape[0]
        return dict(
            poses=f(self.poses)[None], patches=f(self.patches).view(1, NM, 3, 1, 1),
            patches_monodisp=f(self.monodisp).view(1, NM, 1), intrinsics=f(self.intrinsics)[None],
            targets_2d=f(self.targets)[None], weights=f(self.weights)[None],
            ii=g(self.ii), jj=g(self.jj), kk=g(self.kk))


CONFIGS = {
    # name: (n_kf, tracks_per_kf, deg, width, height, iters)
    "cfg1": (8, 64, 8, 640, 480, 3),          # 512 tracks, 4 096 edges  (BASELINE.json configs[0])
    "cfg3": (256, 256, 19, 1024, 436, 10),    # 65 536 tracks, 1 245 184 edges (configs[2], headline)
    "cfg5": (1024, 256, 19, 1024, 436, 10),   # 262 144 tracks, 4 980 736 edges (configs[4])
    "mid": (64, 256, 19, 1024, 436, 10),      # 16 384 tracks, 311 296 edges
    "tiny": (5, 16, 4, 640, 480, 2),
}


def make_window_problem(n_kf, tracks_per_kf, deg, width=1024, height=436, seed=0,
                        fixedp=1, name="", kf_lo=0, kf_hi=None):
    """Sliding-window graph of SURVEY.md §8(d).

    `kf_lo:kf_hi` selects the source keyframes whose tracks (and all their edges) are emitted —
    the keyframe-window shard of SURVEY.md §8(e). Random draws are made for the WHOLE graph
    first, so a shard is bit-identical to the corresponding slice of the full problem.
    """
    rng = np.random.default_rng(seed)
    N, M = n_kf, tracks_per_kf
    deg = min(deg, N)
    intr = np.tile(np.array([500.0, 500.0, width / 2.0, height / 2.0]), (N, 1))

    xi = rng.normal(size=(N, 6)) * np.array([0.01] * 3 + [0.002] * 3)
    xi[0] = 0.0
    gt = _exp(np.cumsum(xi, axis=0))

    x = rng.uniform(100.0, width - 100.0, size=N * M)
    y = rng.uniform(100.0, height - 100.0, size=N * M)
    d = rng.uniform(0.2, 1.2, size=N * M)
    xyd_gt = np.stack([x, y, d], 1)

    src = np.repeat(np.arange(N), M)                      # frame of each patch
    start = np.clip(src - deg // 2, 0, N - deg)
    kk = np.repeat(np.arange(N * M), deg)
    jj = (start[:, None] + np.arange(deg)[None, :]).reshape(-1)
    ii = src[kk]

    tgt, _ = _reproject(gt, xyd_gt, intr, ii, jj, kk)
    tgt = tgt + rng.normal(size=tgt.shape) * 0.5

    noise = rng.normal(size=(N, 6)) * 0.002
    noise[0] = 0.0
    init = _mul(gt, _exp(noise))
    init[0] = gt[0]
    disp0 = d * (1.0 + rng.normal(size=d.shape) * 0.05)
    xyd0 = np.stack([x, y, disp0], 1)

    if kf_hi is None:
        kf_hi = N
    if kf_lo != 0 or kf_hi != N:
        keep = (ii >= kf_lo) & (ii < kf_hi)
        ii, jj, kk, tgt = ii[keep], jj[keep], kk[keep], tgt[keep]

    return BAProblem(
        poses=init.astype(np.float32), patches=xyd0.astype(np.float32), monodisp=d.astype(np.float32),
        intrinsics=intr.astype(np.float32), targets=tgt.astype(np.float32),
        weights=np.ones((ii.shape[0], 2), np.float32),
        ii=ii.astype(np.int64), jj=jj.astype(np.int64), kk=kk.astype(np.int64),
        bounds=[0, 0, width, height], fixedp=fixedp, name=name,
        gt_poses=gt, gt_disp=d)


def make_config(name, seed=0, **kw):
    n_kf, m, deg, w, h, _ = CONFIGS[name]
    return make_window_problem(n_kf, m, deg, w, h, seed=seed, name=name, **kw)


def config_iters(name):
    return CONFIGS[name][5]


def make_slam_problem(n_frames=21, patches_per_frame=96, s_slam=12, kf_stride=2, opt_window=15,
                      removal_window=20, width=960, height=540, seed=0, dyn_frac=0.2,
                      buffer_size=None, name="slam"):
    """Replay of the reference's own graph bookkeeping on synthetic tracks (cfg2 / cfg4 stand-in,
    because DAVIS / Sintel data and the tracker checkpoint are absent).

    Mirrors main/batrack.py: at frame count n (every `kf_stride`-th step, :990) edges are appended
    between the patches of frames n-S_slam..n step kf_stride and frames [n-S_slam, n) (:399-410,
    189-204) — so (patch, frame) pairs repeat across steps with fresh targets, self-edges exist —
    edges whose source frame left the removal window are dropped (:1023-1026), the first
    `fixedp = n - OPTIMIZATION_WINDOW` poses are held fixed (:858-859), the pose buffer is longer
    than the live window (unused tail rows are identity), out-of-frame / invisible observations
    carry zero weight (:773-778) and "dynamic" tracks have zero pose-weight (:790-792).
    Returns (problem with weights = weights_pose, weights_all) like the two calls of update().
    """
    rng = np.random.default_rng(seed)
    n, M = n_frames, patches_per_frame
    N = buffer_size or (n + 3)
    intr = np.tile(np.array([0.9 * width, 0.9 * width, width / 2.0, height / 2.0]), (N, 1))
    xi = rng.normal(size=(N, 6)) * np.array([0.02] * 3 + [0.004] * 3)
    xi[0] = 0.0
    xi[n:] = 0.0
    gt = _exp(np.cumsum(xi, axis=0))
    gt[n:] = np.array([0, 0, 0, 0, 0, 0, 1.0])
    x = rng.uniform(30.0, width - 30.0, size=N * M)
    y = rng.uniform(30.0, height - 30.0, size=N * M)
    d = rng.uniform(0.1, 1.5, size=N * M)
    xyd_gt = np.stack([x, y, d], 1)
    src_of = np.repeat(np.arange(N), M)

    II, JJ, KK = [], [], []
    for step in range(2, n + 1):
        if (step - 1) % kf_stride != 0:
            continue
        lo = max(step - s_slam, 0)
        kf = np.arange(lo, step, kf_stride)
        pk = (kf[:, None] * M + np.arange(M)[None, :]).reshape(-1)
        fr = np.arange(lo, step)
        k_new = np.repeat(pk, fr.shape[0])
        j_new = np.tile(fr, pk.shape[0])
        II.append(src_of[k_new]); JJ.append(j_new); KK.append(k_new)
    ii, jj, kk = np.concatenate(II), np.concatenate(JJ), np.concatenate(KK)
    keep = ii >= n - removal_window
    ii, jj, kk = ii[keep], jj[keep], kk[keep]

    tgt, z = _reproject(gt, xyd_gt, intr, ii, jj, kk)
    dynamic = rng.uniform(size=N * M) < dyn_frac
    tgt = tgt + rng.normal(size=tgt.shape) * 0.7
    tgt[dynamic[kk]] += rng.normal(size=(int(dynamic[kk].sum()), 2)) * 6.0     # moving points
    outlier = rng.uniform(size=ii.shape[0]) < 0.02
    tgt[outlier] += rng.normal(size=(int(outlier.sum()), 2)) * 40.0            # huber territory
    pad = 20
    vis = (tgt[:, 0] >= pad) & (tgt[:, 0] < width - pad) & (tgt[:, 1] >= pad) & (tgt[:, 1] < height - pad)
    vis &= rng.uniform(size=ii.shape[0]) > 0.05
    w_all = np.repeat(vis[:, None].astype(np.float64), 2, 1)
    w_pose = w_all.copy()
    w_pose[dynamic[kk]] = 0.0

    noise = rng.normal(size=(N, 6)) * 0.004
    fixedp = max(n - opt_window, 1)
    noise[:fixedp] = 0.0
    noise[n:] = 0.0
    init = _mul(gt, _exp(noise))
    disp0 = d * (1.0 + rng.normal(size=d.shape) * 0.08)
    mono = d * (1.0 + rng.normal(size=d.shape) * 0.02)
    mono[rng.uniform(size=d.shape) < 0.1] = 0.0          # missing mono depth -> prior disabled

    prob = BAProblem(
        poses=init.astype(np.float32), patches=np.stack([x, y, disp0], 1).astype(np.float32),
        monodisp=mono.astype(np.float32), intrinsics=intr.astype(np.float32),
        targets=tgt.astype(np.float32), weights=w_pose.astype(np.float32),
        ii=ii.astype(np.int64), jj=jj.astype(np.int64), kk=kk.astype(np.int64),
        bounds=[0, 0, width, height], fixedp=fixedp, name=name, gt_poses=gt, gt_disp=d)
    return prob, w_all.astype(np.float32)


def make_random_problem(n_poses=9, n_patches=40, n_edges=300, seed=0, fixedp=2, width=640, height=480,
                        name="random"):
    """Unstructured graph: arbitrary (ii, jj, kk) triples — ii is NOT a function of kk, duplicate
    edges, self-edges, patches that are never observed, poses beyond max(ii,jj) in the buffer,
    zero and fractional weights. Exercises the general (irregular) code path."""
    rng = np.random.default_rng(seed)
    N = n_poses + 2
    intr = np.tile(np.array([420.0, 410.0, width / 2.0 + 3, height / 2.0 - 2]), (N, 1))
    intr += rng.normal(size=intr.shape) * 2.0
    xi = rng.normal(size=(N, 6)) * np.array([0.03] * 3 + [0.01] * 3)
    xi[0] = 0
    gt = _exp(np.cumsum(xi, axis=0))
    x = rng.uniform(60.0, width - 60.0, size=n_patches)
    y = rng.uniform(60.0, height - 60.0, size=n_patches)
    d = rng.uniform(0.2, 1.0, size=n_patches)
    ii = rng.integers(0, n_poses, size=n_edges)
    jj = rng.integers(0, n_poses, size=n_edges)
    kk = rng.integers(0, n_patches - 3, size=n_edges)      # last 3 patches never observed
    dup = rng.integers(0, n_edges, size=n_edges // 10)
    ii = np.concatenate([ii, ii[dup]]); jj = np.concatenate([jj, jj[dup]]); kk = np.concatenate([kk, kk[dup]])
    tgt, _ = _reproject(gt, np.stack([x, y, d], 1), intr, ii, jj, kk)
    tgt += rng.normal(size=tgt.shape) * 1.0
    w = rng.uniform(0.0, 1.0, size=(ii.shape[0], 2))
    w[rng.uniform(size=ii.shape[0]) < 0.1] = 0.0
    noise = rng.normal(size=(N, 6)) * 0.003
    noise[:fixedp] = 0
    init = _mul(gt, _exp(noise))
    disp0 = d * (1.0 + rng.normal(size=d.shape) * 0.05)
    return BAProblem(
        poses=init.astype(np.float32), patches=np.stack([x, y, disp0], 1).astype(np.float32),
        monodisp=d.astype(np.float32), intrinsics=intr.astype(np.float32),
        targets=tgt.astype(np.float32), weights=w.astype(np.float32),
        ii=ii.astype(np.int64), jj=jj.astype(np.int64), kk=kk.astype(np.int64),
        bounds=[0, 0, width, height], fixedp=fixedp, name=name, gt_poses=gt, gt_disp=d)
