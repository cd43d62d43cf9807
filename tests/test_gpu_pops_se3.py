"""GPU tests of the surface next to the BA operator: every form of pops.transform (projective_ops.py:54-105) against the
reference's own Jacobians (tests/golden/cfg1_transform.npz), point_cloud / back_proj / proj_to_frames against the oracle
restatement, and the four forward identities of the reference's lietorch/run_tests.py:16-52 run in fp64 at atol 1e-8
against the CUDA SE3 ops (the reference dispatches float and double, lietorch/include/dispatch.h:37-45)."""
import numpy as np
import pytest
import torch

from conftest import Fixture, rel_err

pytestmark = pytest.mark.gpu


def test_transform_jacobian_and_depth_match_reference():
    from batrack_b200 import projective_ops as pops
    from batrack_b200.lietorch import SE3
    fx, tr = Fixture("cfg1_rgbd"), Fixture("cfg1_transform")
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long().cuda()
    NM = fx.patches.shape[0]
    poses, patches, K = SE3(c(fx.poses)[None]), c(fx.patches).view(1, NM, 3, 1, 1), c(fx.intrinsics)[None]
    ii, jj, kk = g(fx.ii), g(fx.jj), g(fx.kk)
    coords, v, (Ji, Jj, Jz) = pops.transform(poses, patches, K, ii, jj, kk, jacobian=True)
    E = ii.numel()
    assert coords.shape == (1, E, 1, 1, 2) and v.shape == (1, E) and Ji.shape == (1, E, 2, 6) and Jz.shape == (1, E, 2, 1)
    assert rel_err(coords[0, :, 0, 0].cpu().numpy(), tr.coords) < 1e-6
    assert np.array_equal(v[0].cpu().numpy(), tr.valid)
    for ours, ref, name in ((Ji, tr.Ji, "Ji"), (Jj, tr.Jj, "Jj"), (Jz, tr.Jz, "Jz")):
        e = rel_err(ours[0].cpu().numpy(), ref)
        assert e < 5e-6, (name, e)
    # depth=True: third channel = inverse depth in frame j (projective_ops.py:47-50); valid=True: [1,E,1,1] mask
    cd, vd = pops.transform(poses, patches, K, ii, jj, kk, depth=True, valid=True)
    assert cd.shape == (1, E, 1, 1, 3) and vd.shape == (1, E, 1, 1)
    assert torch.equal(cd[..., :2], coords) and torch.equal(vd[:, :, 0, 0], v)
    X1 = (poses[:, jj] * poses[:, ii].inv())[:, :, None, None] * pops.iproj(patches[:, kk], K[:, ii])
    ref = pops.proj(X1, K[:, jj], depth=True)
    assert rel_err(cd.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    # the plain form, and the PRINT=True residual of BA (ba.py:244-245) that is built on it
    assert torch.equal(pops.transform(poses, patches, K, ii, jj, kk), coords)


def test_print_path_reports_the_reference_residual(capsys):
    """BA_rgbd_droid(PRINT=True) prints the mean masked residual norm of ba.py:244-245."""
    from batrack_b200 import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from gpu_util import as_cuda
    from oracle import ba_oracle
    prob = synth.make_config("cfg1")
    t = as_cuda(prob)
    w = torch.ones(1, prob.E, 2, device="cuda")
    BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None, w, prob.lmbda,
                  t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, PRINT=True, fixedp=prob.fixedp, loss=prob.loss,
                  alpha=prob.alpha)
    printed = float(capsys.readouterr().out.strip().splitlines()[-1])
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double()
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long()
    r, _, _, _, _, _, v = ba_oracle.edge_terms(f(prob.poses), f(prob.patches), f(prob.intrinsics), f(prob.targets),
                                               f(prob.weights), g(prob.ii), g(prob.jj), g(prob.kk), prob.bounds, "huber")
    assert abs(printed - float(r.norm(dim=-1).mean())) < 1e-4 * max(1.0, printed)


def test_point_cloud_back_proj_and_proj_to_frames():
    """The point-cloud refresh next to BA (main/batrack.py:440-444, 821-854, 891-893)."""
    from batrack_b200 import projective_ops as pops, synth
    from batrack_b200.lietorch import SE3
    from gpu_util import as_cuda
    from oracle import se3_ops
    ps, _ = synth.make_slam_problem(n_frames=21, patches_per_frame=32, seed=3)
    t = as_cuda(ps)
    NM = ps.patches.shape[0]
    ix = torch.arange(NM, device="cuda") // 32
    pc = pops.point_cloud(SE3(t["poses"]), t["patches"], t["intrinsics"], ix)
    assert pc.shape == (1, NM, 1, 1, 4)
    f = lambda a: torch.from_numpy(a).double()
    P, X, K = f(ps.poses), f(ps.patches), f(ps.intrinsics)
    ixc = ix.cpu()
    x0 = torch.stack([(X[:, 0] - K[ixc, 2]) / K[ixc, 0], (X[:, 1] - K[ixc, 3]) / K[ixc, 1], torch.ones(NM, dtype=torch.float64), X[:, 2]], 1)
    ref = se3_ops.se3_act4(se3_ops.se3_inv(P[ixc].contiguous()), x0)
    assert rel_err(pc[0, :, 0, 0].cpu().numpy(), ref.numpy()) < 2e-6
    # back_proj -> proj_to_frames round trip (projective_ops.py:129-176): a pixel with depth, lifted with c2w and
    # projected back with w2c = c2w^-1 into the same camera, returns to the pixel
    B, n, S = 2, 257, 3
    rng = np.random.default_rng(0)
    xy = torch.from_numpy(rng.uniform(20, 400, (B, n, 2))).float().cuda()
    dep = torch.from_numpy(rng.uniform(0.5, 4.0, (B, n, 1))).float().cuda()
    Kb = torch.tensor([[500.0, 510.0, 320.0, 240.0], [450.0, 455.0, 300.0, 220.0]]).cuda()
    c2w = SE3(t["poses"][0, 1:1 + B].contiguous()).inv().matrix()                   # [B,4,4]
    Pw = pops.back_proj(xy, dep, Kb, c2w)
    Pc = pops.back_proj(xy, dep, Kb)
    assert Pw.shape == (B, n, 4)
    refc = torch.stack([(xy[..., 0] - Kb[:, None, 2]) / Kb[:, None, 0] * dep[..., 0], (xy[..., 1] - Kb[:, None, 3]) / Kb[:, None, 1] * dep[..., 0],
                        dep[..., 0], torch.ones_like(dep[..., 0])], dim=2)
    assert rel_err(Pc.cpu().numpy(), refc.cpu().numpy()) < 1e-6
    assert rel_err(Pw.cpu().numpy(), (c2w @ refc.permute(0, 2, 1)).permute(0, 2, 1).cpu().numpy()) < 2e-6
    w2c = torch.linalg.inv(c2w.double()).float()[:, None].expand(B, S, 4, 4).contiguous()
    Ks = Kb[:, None].expand(B, S, 4).contiguous()
    back = pops.proj_to_frames(Pw, Ks, w2c)
    assert back.shape == (B, S, n, 2)
    assert (back - xy[:, None]).abs().max().item() < 2e-2                              # pixels, fp32 round trip


def test_se3_identities_in_fp64():
    """lietorch/run_tests.py:16-52 (test_exp_log, test_inv, test_adj, test_act), fp64, atol 1e-8, on the CUDA ops."""
    from batrack_b200.lietorch import SE3
    torch.manual_seed(0)
    dev, dt = "cuda", torch.float64
    # exp(log(X)) == X  (run_tests.py:16-21)
    a = 0.2 * torch.randn(2, 3, 4, 6, device=dev, dtype=dt)
    b = SE3.exp(a).log()
    assert b.dtype == dt and torch.allclose(a, b, atol=1e-8)
    # X * X^-1 == identity  (:23-28)
    X = SE3.exp(0.2 * torch.randn(1, 2, 3, 6, device=dev, dtype=dt))
    assert torch.allclose((X * X.inv()).log(), torch.zeros(1, 2, 3, 6, device=dev, dtype=dt), atol=1e-8)
    # X * exp(a) == exp(Ad_X a) * X  (:30-41)
    X = SE3.exp(torch.randn(2, 5, 6, device=dev, dtype=dt))
    a = torch.randn(2, 5, 6, device=dev, dtype=dt)
    c = ((X * SE3.exp(a)) * (SE3.exp(X.adj(a)) * X).inv()).log()
    assert torch.allclose(c, torch.zeros_like(c), atol=1e-8)
    # X * p == matrix(X) @ p  (:44-52)
    X = SE3.exp(torch.randn(1, 6, device=dev, dtype=dt))
    p = torch.randn(1, 3, device=dev, dtype=dt)
    p1 = X.act(p)
    p2 = (X.matrix()[:, :3, :3] @ p[..., None])[..., 0] + X.matrix()[:, :3, 3]
    assert torch.allclose(p1, p2, atol=1e-8)
    # adjT is the transpose of adj: <Ad a, b> == <a, Ad^T b>
    Xs = SE3.exp(torch.randn(7, 6, device=dev, dtype=dt))
    u, w = torch.randn(7, 6, device=dev, dtype=dt), torch.randn(7, 6, device=dev, dtype=dt)
    assert torch.allclose((Xs.adj(u) * w).sum(-1), (u * Xs.adjT(w)).sum(-1), atol=1e-8)
    # fp64 and fp32 kernels agree to fp32 rounding
    assert rel_err(Xs.to(torch.float32).adjT(w.float()).cpu().numpy(), Xs.adjT(w).cpu().numpy()) < 1e-5
