#!/usr/bin/env python
"""Absolute timeline of the band solver's warps within a tile column (solver_trace stamps, SM clock): where each warp
is relative to the moment the factor warp publishes W_J. python tools/solver_timeline.py"""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import synth
from batrack_b200.ba import BA_rgbd_droid
from batrack_b200.lietorch import SE3
from batrack_b200.plan import Plan
prob = synth.make_config("cfg3")
t = {k: v.cuda() for k, v in prob.as_torch().items()}
plan = Plan(t["ii"], t["jj"], t["kk"], prob.poses.shape[0], prob.patches.shape[0])
plan.set_option("solver", "diag"); plan.set_option("solver_trace", 1)
for _ in range(3):
    BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None, t["weights"], prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp, structure_only=False, loss=prob.loss, alpha=prob.alpha, plan=plan)
torch.cuda.synchronize()
tr = plan.read_trace()
h = tr[:16 * 4096].reshape(4096, 16).astype(np.float64)
ncol = 80
base = h[10:ncol, 10]          # W_J published (factor warp, column J)
names = {0: "tile0 top", 1: "tile0 after barP", 2: "tile0 first tile shipped (Dn)", 3: "tile0 end of U", 4: "panel0 top", 5: "panel0 after barA", 6: "panel0 after barW", 7: "panel0 arrive P", 8: "factor top", 9: "factor D loaded", 10: "factor W published"}
print("offsets relative to 'factor W_J published' (column J), mean over columns; same-column stamps and next-column stamps")
for k in range(11):
    print(f"  col J   {names[k]:34s} {np.mean(h[10:ncol, k] - base):8.0f}")
for k in range(11):
    print(f"  col J+1 {names[k]:34s} {np.mean(h[11:ncol + 1, k] - base):8.0f}")
