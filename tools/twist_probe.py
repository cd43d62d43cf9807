#!/usr/bin/env python
"""One twisted band solve (112 key frames) — used under compute-sanitizer to probe the two-CTA path. python tools/twist_probe.py [n_calls] [bad]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import synth
from batrack_b200.ba import BA_rgbd_droid
from batrack_b200.lietorch import SE3
from batrack_b200.plan import Plan
n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 1
bad = len(sys.argv) > 2
prob = synth.make_window_problem(112, 64, 19, seed=3)
t = {k: v.cuda() for k, v in prob.as_torch().items()}
N, NM = t["poses"].shape[1], t["patches"].shape[1]
plan = Plan(t["ii"], t["jj"], t["kk"], N, NM)
for k in range(n_calls):
    ep = -1e9 if (bad and k == 1) else prob.ep
    G, p = BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None, t["weights"],
                         prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=ep, fixedp=prob.fixedp, loss=prob.loss, alpha=prob.alpha, plan=plan)
    torch.cuda.synchronize()
    print("call", k, "ep", ep, "status", plan.status(), "pose checksum", float(G.data.double().sum()), flush=True)
