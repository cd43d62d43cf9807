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
    import synth
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
    from batrack_b200 import projective_ops as pops
    import synth
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


def test_trajectory_handoff_matches_the_reference_recursion():
    """terminate() / get_results() (main/batrack.py:898-915, 1080-1088): every frame's pose through get_pose's recursion
    (:223-228), inverted; one kernel against the same recursion spelled out with the CUDA SE3 ops in fp64
    (themselves pinned on the reference by test_se3_identities_in_fp64 / the se3_ops fixture), including chains of dropped keyframes and the results dictionary's keys."""
    import pickle, tempfile
    from batrack_b200 import lietorch
    from batrack_b200.lietorch import SE3
    from batrack_b200.results import get_results, terminate, trajectory
    rng = np.random.default_rng(5)
    counter, N = 40, 48
    kf_frames = [t for t in range(counter) if t % 3 == 0 or t > 33]          # frames still in the keyframe buffer
    n = len(kf_frames)
    xi = rng.normal(size=(N, 6)) * np.array([0.3] * 3 + [0.2] * 3)
    poses = SE3.exp(torch.from_numpy(xi).float().cuda()).data.clone()
    poses[:, 3:] *= torch.from_numpy(rng.uniform(0.7, 1.4, size=(N, 1))).float().cuda()    # un-normalised quaternions
    tstamps = torch.zeros(N, dtype=torch.int64)
    tstamps[:n] = torch.tensor(kf_frames)
    delta = {}
    for t in range(counter):
        if t not in kf_frames:                                                # dropped: attached to the frame before
            d = SE3.exp(torch.from_numpy(rng.normal(size=(1, 6)) * 0.05).float().cuda()).data
            delta[t] = (t - 1, SE3(d[0]))
    out7, cams = trajectory(poses, tstamps, n, delta, counter)
    # the reference's code path, with the fp64 CUDA ops
    traj = {int(tstamps[i]): poses[i].double() for i in range(n)}

    def get_pose(t):
        if t in traj:
            return SE3(traj[t])
        t0, dP = delta[t]
        return SE3(dP.data.double()) * get_pose(t0)

    ref = lietorch.stack([get_pose(t) for t in range(counter)], dim=0).inv()
    ref7 = ref.data.cpu().numpy()[:, [0, 1, 2, 6, 3, 4, 5]]
    assert np.abs(out7.cpu().numpy() - ref7).max() < 2e-6 * max(1.0, np.abs(ref7).max())
    assert np.abs(cams.cpu().numpy() - ref.matrix().cpu().numpy()).max() < 2e-6 * max(1.0, np.abs(ref7).max())
    p7, ts = terminate(poses, tstamps, n, delta, counter, list(range(counter)))
    assert p7.shape == (counter, 7) and ts.dtype == np.float64
    M, S = 8, 5
    z = lambda *s: torch.zeros(*s).cuda()
    with tempfile.NamedTemporaryFile(suffix=".pkl") as f:
        res = get_results(poses, tstamps, n, delta, counter, list(range(counter)), z(N, 4), torch.ones(N, M).cuda(), z(N, M, S, 3),
                          torch.ones(N, M, S, 1).cuda(), z(N, M, S, 1), z(N, M, S, 1), save_path=f.name)
        back = pickle.load(open(f.name, "rb"))
    assert sorted(res) == sorted(["cams_T_world", "intrinsics", "tstamps", "trajs_2d_disp", "trajs_valid", "trajs_static", "trajs_vis",
                                  "grid_query_frames", "dmaps", "rgbs", "dmaps_gt"])
    assert back["cams_T_world"].shape == (counter, 4, 4) and back["trajs_valid"].shape == (counter, M)
    delta.pop(1)                                                              # frame 1 now has neither pose nor delta
    with pytest.raises(KeyError):
        trajectory(poses, tstamps, n, delta, counter)
