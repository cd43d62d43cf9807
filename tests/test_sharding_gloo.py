"""CPU, world_size 2, gloo: the decomposition the multi-GPU path relies on. Each rank
assembles the reduced camera system of ITS keyframe window's tracks (here with the oracle, on CPU), one
all-reduce sums [S | y], and every rank then solves the same system and back-substitutes its own tracks.
The merged result must equal the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, rel_err  # noqa: F401


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synth
    from oracle import ba_oracle
    n_kf = synth.CONFIGS["cfg1"][0]
    lo, hi = (n_kf * rank) // world, (n_kf * (rank + 1)) // world
    sh = synth.make_config("cfg1", kf_lo=lo, kf_hi=hi)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double()
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long()
    # every rank must lay the reduced system out for the GLOBAL pose count: pad with an edge-free max
    n_total = torch.tensor([int(max(sh.ii.max(), sh.jj.max())) + 1])
    dist.all_reduce(n_total, op=dist.ReduceOp.MAX)
    n = int(n_total) - sh.fixedp
    parts = {}
    ba_oracle.ba_step(f(sh.poses), f(sh.patches), f(sh.monodisp), f(sh.intrinsics), f(sh.targets), f(sh.weights),
                      sh.lmbda, g(sh.ii), g(sh.jj), g(sh.kk), sh.bounds, ep=sh.ep, fixedp=sh.fixedp, loss=sh.loss,
                      alpha=sh.alpha, mode="dense", parts=parts)
    M = 6 * n
    Sy = torch.zeros(M * M + M, dtype=torch.float64)
    nl = parts["n"]
    Sy[:M * M].view(M, M)[:6 * nl, :6 * nl] = parts["S"]
    Sy[M * M:][:6 * nl] = parts["y"]
    dist.all_reduce(Sy)                                   # the one exchange of the sharded path
    S, y = Sy[:M * M].view(M, M), Sy[M * M:]
    dX = ba_oracle._damped_solve(S, y, sh.ep, 1e-4)
    E = parts["E"]                                        # [6 nl, m_local]
    dZ = parts["Q"] * (parts["w"] - E.t() @ dX[:6 * nl])
    disp = f(sh.patches)[:, 2].clone()
    disp[parts["kx"]] += dZ
    own = torch.zeros_like(disp)
    own[parts["kx"]] = 1
    merged = disp * own
    dist.all_reduce(merged)
    if rank == 0:
        full = synth.make_config("cfg1")
        P, D = ba_oracle.run_sequence(full, [full.weights], [False], torch.float64)
        out["dX"] = float((dX.abs().max()))
        out["disp_err"] = float(np.abs(merged.clamp(1e-3, 10).numpy() - D[0]).max() / np.abs(D[0]).max())
        full_parts = {}
        ba_oracle.ba_step(f(full.poses), f(full.patches), f(full.monodisp), f(full.intrinsics), f(full.targets),
                          f(full.weights), full.lmbda, g(full.ii), g(full.jj), g(full.kk), full.bounds, ep=full.ep,
                          fixedp=full.fixedp, loss=full.loss, alpha=full.alpha, parts=full_parts)
        out["S_err"] = float((S - full_parts["S"]).abs().max() / full_parts["S"].abs().max())
        out["dX_err"] = float((dX - full_parts["dX"]).abs().max() / full_parts["dX"].abs().max())
    dist.barrier()
    dist.destroy_process_group()


def test_keyframe_sharding_world2_gloo():
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["S_err"] < 1e-12 and out["dX_err"] < 1e-9 and out["disp_err"] < 1e-9, dict(out)
