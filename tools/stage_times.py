#!/usr/bin/env python
"""Per-stage device times (library CUDA events) of one BA call on a named workload: cfg3 | mid | davis | sintel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth
from batrack_b200.ba import BA_rgbd_droid
from batrack_b200.lietorch import SE3
from batrack_b200.plan import Plan

name = sys.argv[1] if len(sys.argv) > 1 else "davis"
if name == "davis":
    prob, w_all = synth.make_slam_problem(n_frames=25, patches_per_frame=400, seed=7, buffer_size=64)
elif name == "sintel":
    prob, w_all = synth.make_slam_problem(n_frames=50, patches_per_frame=256, seed=4, buffer_size=64, opt_window=64,
                                          removal_window=64, width=1024, height=436, name="sintel_like")
else:
    prob = synth.make_config(name); w_all = prob.weights
t = {k: v.cuda() for k, v in prob.as_torch().items()}
plan = Plan(t["ii"], t["jj"], t["kk"], prob.poses.shape[0], prob.patches.shape[0])
i = plan.info
print(f"{name}: E {i.n_edges} tracks {i.n_tracks} groups {i.n_groups} chunks {i.n_chunks} dmax {i.max_degree} Wmax {i.max_slots} bwb {i.block_bandwidth} banded {i.banded} n_total {i.n_total} fixedp {prob.fixedp}")
plan.enable_timing(True)
for so in (False, True):
    acc = {}
    for k in range(25):
        BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                      t["weights"], prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp,
                      structure_only=so, loss=prob.loss, alpha=prob.alpha, plan=plan)
        tm = plan.last_timing()
        if k >= 5:
            for a, b in tm.items(): acc[a] = acc.get(a, 0) + b / 20
    print("structure_only" if so else "pose+depth   ", {a: round(b * 1e3, 1) for a, b in acc.items()}, "sum", round(sum(acc.values()) * 1e3, 1), "us")
