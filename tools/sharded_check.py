#!/usr/bin/env python
"""Keyframe-sharded BA over N ranks (torchrun): every rank owns the tracks of its keyframe window, poses /
intrinsics / patches are replicated, and the only exchange is one NCCL all-reduce of the reduced camera
system [S | y] per call. Checks the result against the fp64 oracle and against the
single-device run of the same graph.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_check.py [config] [iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "mid"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import Plan
    from gpu_util import as_cuda, run_ours

    n_kf = synth.CONFIGS[cfg][0]
    lo, hi = (n_kf * rank) // world, (n_kf * (rank + 1)) // world
    shard = synth.make_config(cfg, kf_lo=lo, kf_hi=hi)
    t = as_cuda(shard, dev)
    plan = Plan(t["ii"], t["jj"], t["kk"], shard.poses.shape[0], shard.patches.shape[0])
    lay = torch.tensor([plan.info.n_total, plan.info.block_bandwidth], device=dev)
    dist.all_reduce(lay, op=dist.ReduceOp.MAX)
    plan.set_layout(int(lay[0]), int(lay[1]))
    w = torch.ones(1, shard.E, 2, device=dev)
    G, p = SE3(t["poses"]), t["patches"]
    for _ in range(iters):
        G, p = BA_rgbd_droid(G, p, t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None, w, shard.lmbda,
                             t["ii"], t["jj"], t["kk"], shard.bounds, ep=shard.ep, fixedp=shard.fixedp,
                             structure_only=False, loss=shard.loss, alpha=shard.alpha, group=dist.group.WORLD, plan=plan)
    # disparities: every rank updated only its own tracks; gather them (end of update())
    disp = p[0, :, 2, 0, 0].clone()
    own = torch.zeros_like(disp)
    own[plan.tracks().long()] = 1.0
    merged = disp * own
    dist.all_reduce(merged)
    cnt = own.clone()
    dist.all_reduce(cnt)
    disp_all = torch.where(cnt > 0, merged, disp)
    poses = G.data[0]
    ref = poses.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, poses), "ranks disagree on the poses"
    if rank == 0:
        from oracle import ba_oracle
        full = synth.make_config(cfg)
        ws, so = [full.weights] * iters, [False] * iters
        P1, D1 = run_ours(full, ws, so, device=dev)
        P64, D64 = ba_oracle.run_sequence(full, ws, so, torch.float64, mode="sparse")
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        ep, ed = rel(poses.cpu().numpy(), P64[-1]), rel(disp_all.cpu().numpy(), D64[-1])
        sp, sd = rel(poses.cpu().numpy(), P1[-1]), rel(disp_all.cpu().numpy(), D1[-1])
        print(f"sharded {cfg} x{world}: vs fp64 oracle poses {ep:.2e} disps {ed:.2e} | vs single device poses {sp:.2e} disps {sd:.2e}")
        assert ep < 1e-4 and ed < 1e-4 and sp < 1e-5 and sd < 1e-4
        print("SHARDED_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
