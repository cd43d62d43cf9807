#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REFERENCE'S OWN Python BA code
UNMODIFIED (imported from /root/reference/main) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Two stand-in modules are placed ahead of the reference on sys.path (oracle/shims/): the
`lietorch_backends` extension cannot be built (Eigen 3.4.0 is not vendored, setup.py:20) and
`torch_scatter` is not installed. Everything else — main/backend/ba.py, projective_ops.py,
lietorch/groups.py, group_ops.py, broadcasting.py — is the reference's code as it lies.

Each fixture stores the inputs (so tests never depend on RNG stability) and, per iteration, the
reference's outputs in float64 (the truth) and float32 (the yardstick: how far the reference's own
working precision is from the truth).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("BATRACK_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), os.path.join(REF, "main"), REF, ROOT]

from backend.ba import BA, BA_rgbd_droid            # noqa: E402  (reference, verbatim)
import backend.projective_ops as pops               # noqa: E402
from backend.lietorch import SE3                    # noqa: E402
import lietorch_backends as shim                    # noqa: E402

import synth                      # noqa: E402


def run_ref(prob, weights_seq, structure_seq, dtype, variant="rgbd", lmbda=None, loss=None):
    """Run a sequence of reference BA calls; returns per-call (poses[N,7], disps[NM])."""
    t = prob.as_torch(dtype=dtype)
    Gs, patches = SE3(t["poses"]), t["patches"]
    lm = prob.lmbda if lmbda is None else lmbda
    if isinstance(lm, np.ndarray):
        lm = torch.from_numpy(lm).to(dtype)
    out_p, out_d = [], []
    for w, so in zip(weights_seq, structure_seq):
        wt = torch.from_numpy(w).to(dtype)[None]
        if variant == "rgbd":
            Gs, patches = BA_rgbd_droid(
                Gs, patches, t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None, wt, lm,
                t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp,
                structure_only=so, loss=loss or prob.loss, alpha=prob.alpha)
        else:
            Gs, patches = BA(
                Gs, patches, t["intrinsics"], t["targets_2d"], wt, lm, t["ii"], t["jj"], t["kk"],
                prob.bounds, ep=prob.ep, fixedp=prob.fixedp, structure_only=so, loss=loss or prob.loss)
        out_p.append(Gs.data[0].numpy().copy())
        out_d.append(patches[0, :, 2, 0, 0].numpy().copy())
    return np.stack(out_p), np.stack(out_d)


def inputs_of(prob):
    return dict(poses=prob.poses, patches=prob.patches, monodisp=prob.monodisp, intrinsics=prob.intrinsics,
                targets=prob.targets, weights=prob.weights, ii=prob.ii, jj=prob.jj, kk=prob.kk,
                bounds=np.array(prob.bounds, np.float64),
                scalars=np.array([prob.fixedp, prob.ep, prob.lmbda, prob.alpha], np.float64))


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def fixture(name, prob, weights_seq, structure_seq, **kw):
    p64, d64 = run_ref(prob, weights_seq, structure_seq, torch.float64, **kw)
    p32, d32 = run_ref(prob, weights_seq, structure_seq, torch.float32, **kw)
    extra = {}
    if isinstance(kw.get("lmbda"), np.ndarray):
        extra["lmbda_vec"] = kw["lmbda"]
    save(name, **inputs_of(prob), weights_seq=np.stack(weights_seq),
         structure_seq=np.array(structure_seq, np.int64),
         ref64_poses=p64, ref64_disps=d64, ref32_poses=p32, ref32_disps=d32, **extra)
    scale = np.abs(d64).max()
    print(f"   fp32-vs-fp64 of the reference itself: poses {np.abs(p32 - p64).max() / np.abs(p64).max():.2e}"
          f"  disps {np.abs(d32 - d64).max() / scale:.2e}")


def main():
    torch.set_num_threads(os.cpu_count())

    # ---- cfg1: 8 KF / 512 tracks / 4096 edges, 3 Gauss-Newton iterations (BASELINE.json configs[0])
    p = synth.make_config("cfg1")
    fixture("cfg1_rgbd", p, [p.weights] * 3, [False] * 3)
    fixture("cfg1_ba", p, [p.weights] * 3, [False] * 3, variant="ba")
    fixture("cfg1_cauchy", p, [p.weights] * 2, [False] * 2, loss="cauchy")
    fixture("cfg1_trivial_so", p, [p.weights] * 2, [True, False], loss="trivial")

    # ---- per-track lambda tensor (ba.py:299-300): lmbda reshaped to C's shape
    m = np.unique(p.kk).shape[0]
    lam = np.linspace(1e-4, 5e-2, m)
    fixture("cfg1_lmbda_tensor", p, [p.weights] * 2, [False] * 2, lmbda=lam)

    # ---- SLAM-shaped window (duplicates, self-edges, fixedp>1, zero weights), the update() pairing
    ps, w_all = synth.make_slam_problem(n_frames=21, patches_per_frame=32, seed=3)
    fixture("slam_dual", ps, [ps.weights, w_all] * 2, [False, True] * 2)

    # ---- unstructured graph (ii not a function of kk)
    pr = synth.make_random_problem(seed=5)
    fixture("random_rgbd", pr, [pr.weights] * 2, [False] * 2)
    pr2 = synth.make_random_problem(n_poses=30, n_patches=300, n_edges=4000, seed=6, fixedp=1)
    fixture("random2_ba", pr2, [pr2.weights] * 2, [False] * 2, variant="ba")

    # ---- fixedp beyond every pose (n == 0 branch, ba.py:316) and heavy rejection by bounds
    pn = synth.make_config("tiny")
    pn.fixedp = 5
    fixture("tiny_all_fixed", pn, [pn.weights], [False])
    pb = synth.make_config("tiny")
    pb.bounds = [200, 150, 440, 330]
    fixture("tiny_bounds", pb, [pb.weights] * 2, [False] * 2)

    # ---- intermediates of pops.transform(jacobian=True) on cfg1 (fp64), projective_ops.py:54-100
    t = p.as_torch(dtype=torch.float64)
    coords, v, (Ji, Jj, Jz) = pops.transform(SE3(t["poses"]), t["patches"], t["intrinsics"],
                                             t["ii"], t["jj"], t["kk"], jacobian=True)
    save("cfg1_transform", coords=coords[0, :, 0, 0].numpy(), valid=v[0].numpy(),
         Ji=Ji[0].numpy(), Jj=Jj[0].numpy(), Jz=Jz[0].numpy())

    # ---- SE3 primitives (fp64) through the reference's own SE3 class
    g = torch.Generator().manual_seed(11)
    B = 64
    a = torch.randn(B, 6, generator=g, dtype=torch.float64) * torch.tensor([1, 1, 1, .5, .5, .5])
    a[:4, 3:] = 0.0                       # Taylor branch (theta < EPS)
    a[4, 3:] = 1e-8
    b = torch.randn(B, 6, generator=g, dtype=torch.float64)
    X, Y = SE3.exp(a), SE3.exp(b)
    Xr = SE3(X.data * torch.cat([torch.ones(B, 3), torch.full((B, 4), 1.7)], 1).double())  # un-normalised quats
    p4 = torch.randn(B, 4, generator=g, dtype=torch.float64)
    c = torch.randn(B, 6, generator=g, dtype=torch.float64)
    save("se3_ops", a=a.numpy(), b=b.numpy(), p4=p4.numpy(), c=c.numpy(),
         X=X.data.numpy(), Y=Y.data.numpy(), Xr=Xr.data.numpy(),
         inv=Xr.inv().data.numpy(), mul=(Xr * Y).data.numpy(), act4=Xr.act(p4).numpy(),
         adjT=Xr.adjT(c).numpy(), adj=Xr.adj(c).numpy(), log=X.log().numpy(),
         retr=Y.retr(a).data.numpy(), matrix=Xr.matrix().numpy())


if __name__ == "__main__":
    main()
