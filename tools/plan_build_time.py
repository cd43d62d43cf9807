#!/usr/bin/env python
"""Cost of a topology change: ba_plan_create (exact plan: allocations + two read-backs) against ba_plan_update on a
capacity plan (device-side re-derivation, no synchronisation, no allocation): cfg3 | davis | mid | sintel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth
from batrack_b200.plan import CapacityPlan, Plan

for name in sys.argv[1:] or ["davis", "cfg3"]:
    if name == "davis":
        prob, _ = synth.make_slam_problem(n_frames=25, patches_per_frame=400, seed=7, buffer_size=64)
    elif name == "sintel":
        prob, _ = synth.make_slam_problem(n_frames=50, patches_per_frame=256, seed=4, buffer_size=64, opt_window=64, removal_window=64, width=1024, height=436)
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
        info = p.info
        del p
    cp = CapacityPlan(N, NM, cap_edges=prob.E, cap_groups=info.n_groups + 8, cap_pattern=info.n_groups * info.max_degree + 64)
    host, dev = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        cp.update(t["ii"], t["jj"], t["kk"])
        e1.record()
        host.append(1e3 * (time.perf_counter() - t0))
        torch.cuda.synchronize()
        dev.append(e0.elapsed_time(e1))
        cp.finalize()
    md = lambda x: sorted(x[2:])[len(x[2:]) // 2]
    print(f"{name}: E {prob.E}  ba_plan_create ms: first {ts[0]:.2f} median {md(ts):.3f} | ba_plan_update: host call {md(host):.3f} ms (no sync), "
          f"device {md(dev):.3f} ms", flush=True)
