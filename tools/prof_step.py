#!/usr/bin/env python
"""A few BA steps on a workload for runs under ncu: python tools/prof_step.py cfg3 [key=value ...] (plan options)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
plan = Plan(t["ii"], t["jj"], t["kk"], prob.poses.shape[0], prob.patches.shape[0])
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    plan.set_option(k, v if k == "solver" and not v.isdigit() else int(v))
for _ in range(4):
    BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                  t["weights"], prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp,
                  structure_only=False, loss=prob.loss, alpha=prob.alpha, plan=plan)
torch.cuda.synchronize()
