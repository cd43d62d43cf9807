#!/usr/bin/env python
"""Phase clock stamps of the tensor-core Schur kernel (solver_trace bit 1): python tools/schur_trace.py cfg3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synth
from batrack_b200.ba import BA_rgbd_droid
from batrack_b200.lietorch import SE3
from batrack_b200.plan import Plan
prob = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
t = {k: v.cuda() for k, v in prob.as_torch().items()}
plan = Plan(t["ii"], t["jj"], t["kk"], prob.poses.shape[0], prob.patches.shape[0])
plan.set_option("solver_trace", 2)
for _ in range(1):
    BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                  t["weights"], prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp,
                  structure_only=False, loss=prob.loss, alpha=prob.alpha, plan=plan)
torch.cuda.synchronize()
tr = plan.read_trace()[:16 * 4096].reshape(4096, 16)[:plan.info.n_groups]
a = tr[:, :8].astype(np.float64); b = tr[:, 8:].astype(np.float64)
t0 = a[:, 0].min()
names = ["setup", "convert loop", "wait done", "flush", "final barrier"]
for k in range(5):
    d = a[:, k + 1] - a[:, k]
    print(f"conv warp 0  {names[k]:14s} mean {d.mean():8.0f} min {d.min():8.0f} max {d.max():8.0f} cycles")
print("MMA warp: start->ready", (b[:, 1] - b[:, 0]).mean(), " issue loop", (b[:, 2] - b[:, 1]).mean())
#print("per chunk: cp.async wait", (tr[:,6]/8).mean(), " LDS", (tr[:,15]/8).mean(), " empty wait", (tr[:,7]/8).mean(), " convert+store+fence+arrive", (tr[:,14]/8).mean())
print("CTA start spread", (a[:, 0] - t0).max(), " kernel span", (a[:, 5] - t0).max())
