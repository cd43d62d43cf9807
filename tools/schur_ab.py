#!/usr/bin/env python
"""A/B of the Schur kernels (tcgen05 3xTF32 vs SIMT fp32) on one workload: per-stage device times with / without the
streaming hand-over, and agreement of S, y, dX and the outputs (relative to max-abs) with the SIMT kernel's."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synth
from batrack_b200.ba import BA_rgbd_droid
from batrack_b200.lietorch import SE3
from batrack_b200.plan import Plan

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
if name == "davis":
    prob, _ = synth.make_slam_problem(n_frames=25, patches_per_frame=400, seed=7, buffer_size=64)
else:
    prob = synth.make_config(name)
t = {k: v.cuda() for k, v in prob.as_torch().items()}
N, NM = prob.poses.shape[0], prob.patches.shape[0]
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def call(plan):
    return BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                         t["weights"], prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp,
                         structure_only=False, loss=prob.loss, alpha=prob.alpha, plan=plan)


ref = None
combos = [(1, 0, 64, 4), (0, 0, 64, 4)]
for schur, stream, tu, acc_chunks in combos:
    if True:
        os.environ["BA_STREAM_TU"] = str(tu)
        plan = Plan(t["ii"], t["jj"], t["kk"], N, NM)
        plan.set_option("schur", schur)
        plan.set_option("stream", stream)
        plan.set_option("schur_acc", acc_chunks)
        n = plan.info.n_total - prob.fixedp
        plan.enable_timing(True)
        acc = {}
        for k in range(25):
            G, p = call(plan)
            tm = plan.last_timing()
            if k >= 5:
                for a, b in tm.items():
                    acc[a] = acc.get(a, 0) + b / 20
        torch.cuda.synchronize()
        dbg = {k: v.double().cpu().numpy() for k, v in plan.debug(n).items()}
        out = dict(dbg, poses=G.data.double().cpu().numpy(), disps=p[0, :, 2, 0, 0].double().cpu().numpy())
        if ref is None:
            ref = out
        errs = " ".join(f"{k} {rel(out[k], ref[k]):.1e}" for k in ("S", "y", "dX", "poses", "disps"))
        print(f"{name} schur={'simt' if schur else 'tc'} stream={stream} stream_tu={tu} acc={32 * acc_chunks}: " + " ".join(f"{a} {b * 1e3:.1f}" for a, b in acc.items()) +
              f" | sum {sum(acc.values()) * 1e3:.1f} us | status {plan.status()} | vs simt: {errs}", flush=True)
        del plan
