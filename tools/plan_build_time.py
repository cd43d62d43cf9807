#!/usr/bin/env python
"""Wall time of building a topology plan (what the first BA call on a changed graph pays): cfg3 | davis | mid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batrack_b200 import synth
from batrack_b200.plan import Plan

for name in sys.argv[1:] or ["davis", "cfg3"]:
    if name == "davis":
        prob, _ = synth.make_slam_problem(n_frames=25, patches_per_frame=400, seed=7, buffer_size=64)
    else:
        prob = synth.make_config(name)
    t = {k: v.cuda() for k, v in prob.as_torch().items()}
    N, NM = prob.poses.shape[0], prob.patches.shape[0]
    ts = []
    for k in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p = Plan(t["ii"], t["jj"], t["kk"], N, NM)
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
        del p
    print(f"{name}: E {prob.E}  plan build ms: first {ts[0]:.2f}  median of rest {sorted(ts[1:])[len(ts)//2]:.3f}  min {min(ts):.3f}")
