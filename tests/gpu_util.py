"""Helpers shared by the GPU parity tests: run the product path (batrack_b200 -> C ABI -> CUDA) on a
fixture / synthetic problem the same way tests/golden/make_golden.py ran the reference."""
import numpy as np
import torch


def as_cuda(prob, device="cuda:0"):
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=torch.float32)
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=torch.int64)
    NM = prob.patches.shape[0]
    return dict(poses=f(prob.poses)[None], patches=f(prob.patches).view(1, NM, 3, 1, 1),
                patches_monodisp=f(prob.monodisp).view(1, NM, 1), intrinsics=f(prob.intrinsics)[None],
                targets_2d=f(prob.targets)[None], ii=g(prob.ii), jj=g(prob.jj), kk=g(prob.kk))


def run_ours(prob, weights_seq, structure_seq, variant="rgbd", lmbda=None, loss=None, device="cuda:0",
             return_plan=False):
    from batrack_b200.ba import BA, BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import get_plan
    t = as_cuda(prob, device)
    Gs, patches = SE3(t["poses"]), t["patches"]
    lm = prob.lmbda if lmbda is None else lmbda
    if isinstance(lm, np.ndarray):
        lm = torch.from_numpy(lm).to(device=device, dtype=torch.float32)
    P, D = [], []
    for w, so in zip(weights_seq, structure_seq):
        wt = torch.from_numpy(np.ascontiguousarray(w)).to(device=device, dtype=torch.float32)[None]
        if variant == "rgbd":
            Gs, patches = BA_rgbd_droid(Gs, patches, t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                                        wt, lm, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep,
                                        fixedp=prob.fixedp, structure_only=bool(so), loss=loss or prob.loss,
                                        alpha=prob.alpha)
        else:
            Gs, patches = BA(Gs, patches, t["intrinsics"], t["targets_2d"], wt, lm, t["ii"], t["jj"], t["kk"],
                             prob.bounds, ep=prob.ep, fixedp=prob.fixedp, structure_only=bool(so),
                             loss=loss or prob.loss)
        P.append(Gs.data[0].cpu().numpy().copy())
        D.append(patches[0, :, 2, 0, 0].cpu().numpy().copy())
    if return_plan:
        return np.stack(P), np.stack(D), get_plan(t["ii"], t["jj"], t["kk"], t["poses"].shape[1], t["patches"].shape[1]), t
    return np.stack(P), np.stack(D)
