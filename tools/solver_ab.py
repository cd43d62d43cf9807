#!/usr/bin/env python
"""A/B of the reduced-system solvers on one workload (cfg3 | mid | davis | cfg5): per-stage device times with and
without the Schur -> solve streaming hand-over, agreement of dX between the solvers, and the per-column phase
means of the band solver's trace (SM cycles)."""
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
solvers = sys.argv[2].split(",") if len(sys.argv) > 2 else ["diag", "mma"]
if name == "sintel":
    prob, _ = synth.make_slam_problem(n_frames=50, patches_per_frame=256, seed=4, buffer_size=64, opt_window=64, removal_window=64, width=1024, height=436, name="sintel_like")
elif name == "davis":
    prob, _ = synth.make_slam_problem(n_frames=25, patches_per_frame=400, seed=7, buffer_size=64)
else:
    prob = synth.make_config(name)
t = {k: v.cuda() for k, v in prob.as_torch().items()}
N, NM = prob.poses.shape[0], prob.patches.shape[0]


def call(plan):
    return BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                         t["weights"], prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp,
                         structure_only=False, loss=prob.loss, alpha=prob.alpha, plan=plan)


ref = None
for sv in solvers:
    for stream in (0,):
        plan = Plan(t["ii"], t["jj"], t["kk"], N, NM)
        plan.set_option("solver", sv)
        plan.set_option("stream", stream)
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
        dX = plan.debug(n)["dX"].double().cpu().numpy()
        st = plan.status()
        if ref is None:
            ref = dX
        err = np.abs(dX - ref).max() / max(np.abs(ref).max(), 1e-30)
        print(f"{name} solver={sv} stream={stream}: " + " ".join(f"{a} {b * 1e3:.1f}" for a, b in acc.items()) +
              f" | sum {sum(acc.values()) * 1e3:.1f} us | status {st} | dX vs first {err:.2e}", flush=True)
        if sv == "tiles":
            plan.enable_timing(False)
            plan.set_option("solver_trace", 4)
            call(plan)
            nt = (6 * n + 7) // 8
            h = plan.read_trace()[:16 * 4096].reshape(4096, 16)[1:nt - 2].astype(np.float64)
            m = lambda x: float(x.mean())
            print(f"   tiles trace ({nt} tile columns): [P] {m(h[:, 1] - h[:, 0]):.0f} barrier {m(h[:, 2] - h[:, 1]):.0f} | warp 0: update {m(h[:, 3] - h[:, 2]):.0f} factor {m(h[:, 4] - h[:, 3]):.0f} "
                  f"wait {m(h[:, 5] - h[:, 4]):.0f} | warp 5: pairs {m(h[:, 9] - h[:, 8]):.0f} | warp 4: row load {m(h[:, 13] - h[:, 12]):.0f} | step {m(h[1:, 0] - h[:-1, 0]):.0f}")
        if stream == 0 and sv in ("diag", "mma"):
            plan.enable_timing(False)
            plan.set_option("solver_trace", 1)
            call(plan)
            tr = plan.read_trace()
            M = 6 * n
            nt = (M + 7) // 8
            twist = nt >= plan.get_option("twist_min")
            ncol = (nt - 16) // 2 if twist else nt
            h = tr[:16 * 4096].reshape(4096, 16)
            if ncol > 4 and sv == "diag":
                a, nx = h[1:ncol - 1], h[2:ncol]
                m = lambda x: float(x.mean())
                step = (h[ncol - 1, 0] - h[1, 0]) / (ncol - 2)
                print(f"   trace: tiles {nt} twist {int(twist)} | tile warp 0: fetch+wait panels {m(a[:, 1] - a[:, 0]):.0f} first tiles+ship {m(a[:, 2] - a[:, 1]):.0f} "
                      f"rest of U {m(a[:, 3] - a[:, 2]):.0f} loop {m(nx[:, 0] - a[:, 3]):.0f} | panel warp 0: wait A {m(a[:, 5] - a[:, 4]):.0f} wait W {m(a[:, 6] - a[:, 5]):.0f} "
                      f"tiles {m(a[:, 7] - a[:, 6]):.0f} deferred {m(nx[:, 4] - a[:, 7]):.0f} | factor warp: wait D {m(a[:, 9] - a[:, 8]):.0f} A {m(a[:, 10] - a[:, 9]):.0f} "
                      f"loop {m(nx[:, 8] - a[:, 10]):.0f} | step {step:.0f} cycles")
            elif ncol > 4:
                a, nx = h[1:ncol - 1], h[2:ncol]
                d = [(a[:, 1] - a[:, 0]).mean(), (a[:, 2] - a[:, 1]).mean(), (a[:, 3] - a[:, 2]).mean(),
                     (a[:, 4] - a[:, 3]).mean(), (a[:, 5] - a[:, 4]).mean(), (nx[:, 0] - a[:, 5]).mean(),
                     (a[:, 9] - a[:, 8]).mean(), (a[:, 10] - a[:, 9]).mean(), (nx[:, 8] - a[:, 10]).mean()]
                step = (h[ncol - 1, 0] - h[1, 0]) / (ncol - 2)
                print(f"   trace: tiles {nt} twist {int(twist)} | tile warp 0: top {d[0]:.0f} waitW {d[1]:.0f} P {d[2]:.0f} waitP {d[3]:.0f} "
                      f"U {d[4]:.0f} tail {d[5]:.0f} | factor warp: waitD {d[6]:.0f} A {d[7]:.0f} tail {d[8]:.0f} | step {step:.0f} cycles")
            for sd in range(2 if twist else 1):
                ph = tr[16 * 4096 + sd * 8: 16 * 4096 + sd * 8 + 8]
                print(f"   side {sd}: factor+forward {ph[1] - ph[0]}  back-substitution {ph[2] - ph[1]}  tail {ph[3] - ph[2]} cycles | chain warp waited {ph[4]} for the far field, {ph[5]} for rows; helper warp 0 waited {ph[7]} for x")
        del plan
