"""GPU parity tests proper: the CUDA path (through the C ABI) against the fp64 golden vectors of the
reference's own code and against the oracle restatement on seeded inputs. Tolerance: 1e-4 relative
(max-abs-diff / max-abs), the bar BASELINE.json states; the reference's own fp32 error against the same
fp64 truth is printed next to ours."""
import numpy as np
import pytest
import torch

from conftest import BA_FIXTURES, Fixture, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _golden_large(name):
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    return np.load(path)


def _oracle():
    from oracle import ba_oracle
    return ba_oracle


@pytest.mark.parametrize("name", sorted(BA_FIXTURES))
def test_ba_sequence_matches_reference_fp64(name):
    from gpu_util import run_ours
    fx = Fixture(name)
    variant, loss = BA_FIXTURES[name]
    lm = fx.lmbda_vec if hasattr(fx, "lmbda_vec") else None
    P, D = run_ours(fx, fx.weights_seq, fx.structure_seq, variant=variant, lmbda=lm, loss=loss)
    ep, ed = rel_err(P, fx.ref64_poses), rel_err(D, fx.ref64_disps)
    rp, rd = rel_err(fx.ref32_poses, fx.ref64_poses), rel_err(fx.ref32_disps, fx.ref64_disps)
    print(f"\n{name}: ours vs fp64 poses {ep:.2e} disps {ed:.2e} | reference fp32 vs fp64 poses {rp:.2e} disps {rd:.2e}")
    assert np.isfinite(P).all() and np.isfinite(D).all()
    assert ep < TOL, f"poses {ep}"
    assert ed < TOL, f"disps {ed} (reference fp32 itself: {rd})"


@pytest.mark.parametrize("name", ["cfg1_rgbd", "slam_dual", "random_rgbd", "random2_ba", "tiny_bounds"])
def test_stagewise_against_oracle(name):
    """First call only: reduced system S, y, per-track Q, w, pose update dX and depth update dZ."""
    from gpu_util import run_ours
    fx = Fixture(name)
    variant, loss = BA_FIXTURES[name]
    P, D, plan, t = run_ours(fx, fx.weights_seq[:1], [False], variant=variant, loss=loss, return_plan=True)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double()
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long()
    parts = {}
    _oracle().ba_step(f(fx.poses), f(fx.patches), f(fx.monodisp) if variant == "rgbd" else None, f(fx.intrinsics),
                      f(fx.targets), f(fx.weights_seq[0]), fx.lmbda, g(fx.ii), g(fx.jj), g(fx.kk), fx.bounds,
                      ep=fx.ep, fixedp=fx.fixedp, structure_only=False, loss=loss or fx.loss, alpha=fx.alpha,
                      parts=parts)
    n = parts["n"]
    assert plan.info.n_tracks == parts["m"]
    assert np.array_equal(plan.tracks().cpu().numpy(), parts["kx"].numpy())          # == torch.unique(kk), bit-exact
    dbg = {k: v.cpu().numpy() for k, v in plan.debug(n).items()}
    errs = {k: rel_err(dbg[k], parts[k].numpy()) for k in ("S", "y", "Q", "w", "dX", "dZ")}
    print(f"\n{name}: " + "  ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert errs["S"] < 2e-5 and errs["y"] < 2e-5 and errs["Q"] < 1e-5 and errs["w"] < 1e-4   # w sums residuals: ~eps*|pixel|/|r|
    assert errs["dX"] < TOL and errs["dZ"] < 5 * TOL
    assert plan.status() == 0


def test_plan_structure_cfg1():
    import synth
    from batrack_b200.plan import Plan
    prob = synth.make_config("cfg1")
    g = lambda a: torch.from_numpy(a).cuda()
    plan = Plan(g(prob.ii), g(prob.jj), g(prob.kk), prob.poses.shape[0], prob.patches.shape[0])
    i = plan.info
    assert (i.n_edges, i.n_total, i.n_tracks) == (4096, 8, 512)
    assert i.n_groups == 8 and i.max_degree == 8 and i.max_slots == 8 and i.perm_identity == 1
    assert i.block_bandwidth == 7


def test_plan_rejects_bad_indices():
    from batrack_b200.plan import Plan
    ii = torch.tensor([0, 1, 9], device="cuda")
    with pytest.raises(RuntimeError, match="out of range"):
        Plan(ii, ii.clone(), ii.clone(), 4, 16)


def test_permuted_edges_give_same_answer():
    """The caller's edge order is arbitrary (main/batrack.py appends per keyframe step); a shuffled copy
    of cfg1 must give the same update up to fp32 summation order."""
    from gpu_util import run_ours
    fx = Fixture("cfg1_rgbd")
    P0, D0 = run_ours(fx, fx.weights_seq, fx.structure_seq)
    perm = np.random.default_rng(0).permutation(fx.ii.shape[0])
    for k in ("ii", "jj", "kk", "targets"):
        setattr(fx, k, getattr(fx, k)[perm])
    P1, D1 = run_ours(fx, fx.weights_seq[:, perm], fx.structure_seq)
    assert rel_err(P1, fx.ref64_poses) < TOL and rel_err(D1, fx.ref64_disps) < 2 * TOL
    assert rel_err(P1, P0) < 5e-5 and rel_err(D1, D0) < 1e-4


def test_strided_targets_view():
    """main/batrack.py:871 passes targets_3d[..., :2] (row stride 3)."""
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from gpu_util import as_cuda
    prob = synth.make_config("cfg1")
    t = as_cuda(prob)
    w = torch.ones(1, prob.E, 2, device="cuda")
    t3 = torch.cat([t["targets_2d"], torch.rand(1, prob.E, 1, device="cuda")], dim=-1)
    args = (t["intrinsics"],)
    kw = dict(ep=prob.ep, fixedp=1, loss="huber", alpha=prob.alpha)
    G0, p0 = BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], *args, t["targets_2d"], None, w,
                           1e-4, t["ii"], t["jj"], t["kk"], prob.bounds, **kw)
    G1, p1 = BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], *args, t3[..., :2], t3[..., 2:], w,
                           1e-4, t["ii"], t["jj"], t["kk"], prob.bounds, **kw)
    assert rel_err(G1.data.cpu().numpy(), G0.data.cpu().numpy()) < 2e-6
    assert rel_err(p1.cpu().numpy(), p0.cpu().numpy()) < 2e-5


def test_inputs_untouched_and_outputs_fresh():
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from gpu_util import as_cuda
    prob = synth.make_config("tiny")
    t = as_cuda(prob)
    before = {k: v.clone() for k, v in t.items()}
    w = torch.ones(1, prob.E, 2, device="cuda")
    G, p = BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                         w, 1e-4, t["ii"], t["jj"], t["kk"], prob.bounds, ep=10.0, fixedp=1, loss="huber", alpha=0.05)
    for k, v in t.items():
        assert torch.equal(v, before[k]), k
    assert isinstance(G, SE3) and G.data.shape == t["poses"].shape and p.shape == t["patches"].shape
    assert G.data.data_ptr() != t["poses"].data_ptr() and p.data_ptr() != t["patches"].data_ptr()


def test_contract_errors():
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from gpu_util import as_cuda
    prob = synth.make_config("tiny")
    t = as_cuda(prob)
    w = torch.ones(1, prob.E, 2, device="cuda")
    call = lambda **o: BA_rgbd_droid(SE3(o.get("poses", t["poses"])), t["patches"], t["patches_monodisp"],
                                     t["intrinsics"], t["targets_2d"], None, w, 1e-4, t["ii"], t["jj"], t["kk"],
                                     prob.bounds, ep=10.0, fixedp=1, loss=o.get("loss", "huber"), alpha=0.05)
    with pytest.raises(NotImplementedError):
        call(loss="tukey")                                               # ba.py:98-99
    with pytest.raises(RuntimeError, match="CUDA"):
        call(poses=t["poses"].cpu())                                     # no CPU fallback
    with pytest.raises(TypeError):
        call(poses=t["poses"].double())


def test_cholesky_failure_and_nan_are_silent():
    """ba.py:9-13: a failed factorisation leaves the poses unchanged; depths still move."""
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import get_plan
    from gpu_util import as_cuda
    prob = synth.make_config("tiny")
    t = as_cuda(prob)
    w = torch.ones(1, prob.E, 2, device="cuda")
    G, p = BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                         w, 1e-4, t["ii"], t["jj"], t["kk"], prob.bounds, ep=-1e9, fixedp=1, loss="huber", alpha=0.05)
    plan = get_plan(t["ii"], t["jj"], t["kk"], t["poses"].shape[1], t["patches"].shape[1])
    assert plan.status() & 1
    assert rel_err(G.data.cpu().numpy(), t["poses"].cpu().numpy()) < 1e-6
    assert torch.isfinite(p).all() and not torch.equal(p, t["patches"])


def test_nan_retry_branch_matches_reference_semantics():
    """ba.py:324-325 (rgbd variant): NaN in dX -> one more solve with lm = 1e-3. A NaN target on an edge is rejected by
    the masks (ba.py:233-242) but, exactly as in the reference, the mask multiplies (0 * NaN = NaN): the reduced
    right-hand side turns NaN while S stays finite, the factorisation succeeds, dX is NaN, the retry is taken (status
    bit 1) and — S being the same — ends NaN as well. Outputs must carry NaN exactly where the oracle's do and agree
    elsewhere. (A retry that *repairs* dX needs an overflow inside the fp32 LAPACK solve; the fp64 band solver has no such
    case, so the branch is pinned on what both sides can reach.) BA (no retry in the reference, ba.py:205-207) must not
    set the bit."""
    import synth
    from batrack_b200.ba import BA, BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import get_plan
    from gpu_util import as_cuda
    prob = synth.make_config("cfg1")
    prob.targets = prob.targets.copy()
    prob.targets[37] = np.nan
    t = as_cuda(prob)
    w = torch.ones(1, prob.E, 2, device="cuda")
    plan = get_plan(t["ii"], t["jj"], t["kk"], t["poses"].shape[1], t["patches"].shape[1])
    G, p = BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None, w,
                         prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp, loss=prob.loss,
                         alpha=prob.alpha)
    st = plan.status()
    assert st & 2 and not st & 1, st                                       # retry taken, factorisation itself fine
    P64, D64 = _oracle().run_sequence(prob, [prob.weights], [False], torch.float64)
    Pg, Dg = G.data[0].cpu().numpy(), p[0, :, 2, 0, 0].cpu().numpy()
    assert np.array_equal(np.isnan(Pg), np.isnan(P64[0])) and np.isnan(Pg).any()
    assert np.array_equal(np.isnan(Dg), np.isnan(D64[0]))
    fp, fd = ~np.isnan(P64[0]), ~np.isnan(D64[0])
    assert fp.any() and rel_err(Pg[fp], P64[0][fp]) < TOL                  # the fixed pose stays finite
    assert not fd.any() or rel_err(Dg[fd], D64[0][fd]) < TOL
    BA(SE3(t["poses"]), t["patches"], t["intrinsics"], t["targets_2d"], w, prob.lmbda, t["ii"], t["jj"], t["kk"],
       prob.bounds, ep=prob.ep, fixedp=prob.fixedp, loss=prob.loss)
    assert not plan.status() & 2


def test_streaming_give_up_path_is_correct():
    """The band solver starts next to the Schur kernel and waits for its completion flags with a bounded spin. When the
    producer does not get there in time (kernels serialised by a profiler, a busy GPU) it gives up, reports status bit 3
    and a stand-by launch redoes the solve on the finished system. Forced here on the headline graph with a spin bound of
    a few microseconds (BA_OPT_SPIN_CAP) and a Schur kernel throttled to one CTA per SM (BA_OPT_STREAM_SMEM_KB); the
    result must still match the fp64 oracle."""
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import Plan
    from gpu_util import as_cuda
    prob = synth.make_config("cfg3")
    t = as_cuda(prob)
    plan = Plan(t["ii"], t["jj"], t["kk"], t["poses"].shape[1], t["patches"].shape[1])
    w = torch.from_numpy(prob.weights).cuda()[None]

    def call():
        return BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                             w, prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp,
                             loss=prob.loss, alpha=prob.alpha, plan=plan)

    plan.set_option("stream", 1)
    G0, p0 = call()
    assert plan.status() == 0
    plan.set_option("spin_cap", 1)
    plan.set_option("stream_smem_kb", 200)
    assert plan.get_option("spin_cap") == 1 and plan.get_option("stream") == 1
    G1, p1 = call()
    assert plan.status() & 8, "the streamed solve was expected to give up"
    assert rel_err(G1.data.cpu().numpy(), G0.data.cpu().numpy()) < 1e-6 and rel_err(p1.cpu().numpy(), p0.cpu().numpy()) < 1e-6
    z = _golden_large("cfg3_x10_sparse64.npz")
    assert rel_err(G1.data[0].cpu().numpy(), z["poses"][0]) < TOL
    assert rel_err(p1[0, :, 2, 0, 0].cpu().numpy(), z["disps"][0].astype(np.float64)) < TOL
    plan.set_option("spin_cap", 0)
    plan.set_option("stream_smem_kb", 0)
    call()
    assert plan.status() == 0


def test_se3_ops_match_reference():
    from batrack_b200.lietorch import SE3
    z = Fixture("se3_ops")
    c = lambda a: torch.from_numpy(a).float().cuda()
    X, Y, Xr = SE3(c(z.X)), SE3(c(z.Y)), SE3(c(z.Xr))
    chk = lambda ours, ref, tol=2e-6: rel_err(ours.cpu().numpy(), ref) < tol
    assert chk(SE3.exp(c(z.a)).data, z.X)
    assert chk(Xr.inv().data, z.inv)
    assert chk((Xr * Y).data, z.mul)
    assert chk(Xr.act(c(z.p4)), z.act4)
    assert chk(Xr.adjT(c(z.c)), z.adjT)
    assert chk(Xr.adj(c(z.c)), z.adj)
    assert chk(X.log(), z.log, 1e-5)
    assert chk(Y.retr(c(z.a)).data, z.retr)
    assert chk(Xr.matrix(), z.matrix)
    # broadcasting like projective_ops.py:66 (Gij[:, :, None, None] * X0)
    G = SE3(c(z.X)[None, :, None, None])
    pts = c(z.p4)[None, :, None, None].expand(1, 64, 2, 2, 4)
    assert chk(G * pts, np.broadcast_to(SE3(c(z.X)).act(c(z.p4)).cpu().numpy()[None, :, None, None], (1, 64, 2, 2, 4)), 1e-6)


def test_reproject_matches_oracle():
    from batrack_b200 import projective_ops as pops
    import synth
    from batrack_b200.lietorch import SE3
    from gpu_util import as_cuda
    ps, _ = synth.make_slam_problem(n_frames=21, patches_per_frame=32, seed=3)
    t = as_cuda(ps)
    coords, v = pops.transform(SE3(t["poses"]), t["patches"], t["intrinsics"], t["ii"], t["jj"], t["kk"], valid=True)
    f = lambda a: torch.from_numpy(a).double()
    g = lambda a: torch.from_numpy(a).long()
    c64, v64, *_ = _oracle().reproject_with_jacobians(f(ps.poses), f(ps.patches), f(ps.intrinsics), g(ps.ii), g(ps.jj), g(ps.kk))
    assert rel_err(coords[0, :, 0, 0].cpu().numpy(), c64.numpy()) < 1e-5
    assert np.array_equal(v[0, :, 0, 0].cpu().numpy(), v64.numpy())


def test_mid_graph_against_sparse_oracle():
    """64 keyframes / 16 384 tracks / 311 296 edges, 3 iterations, against the sparse-aware fp64 oracle
    (banded reduced system, n = 63)."""
    import synth
    from gpu_util import run_ours
    prob = synth.make_config("mid")
    ws, so = [prob.weights] * 3, [False] * 3
    P, D = run_ours(prob, ws, so)
    P64, D64 = _oracle().run_sequence(prob, ws, so, torch.float64, mode="sparse")
    ep, ed = rel_err(P, P64), rel_err(D, D64)
    print(f"\nmid: poses {ep:.2e} disps {ed:.2e}")
    assert ep < TOL and ed < TOL


def test_headline_graph_all_ten_iterations():
    """cfg3 (256 KF / 65 536 tracks / 1 245 184 edges, BASELINE.json configs[2]: 10 LM iterations): EVERY one of the 10
    chained iterations against the sparse fp64 oracle's results (tests/golden/cfg3_x10_sparse64.npz, produced by
    tests/golden/make_golden_large.py), plus the size-independent properties."""
    import synth
    from gpu_util import run_ours
    prob = synth.make_config("cfg3")
    z = _golden_large("cfg3_x10_sparse64.npz")
    assert int(z["edges"]) == prob.E and int(z["iters"]) == 10
    P, D = run_ours(prob, [prob.weights] * 10, [False] * 10)
    assert np.isfinite(P).all() and np.isfinite(D).all()
    assert rel_err(P[:, 0], np.broadcast_to(prob.poses[0], P[:, 0].shape)) < 1e-6
    errs_p = [np.abs(P[k] - prob.gt_poses).max() for k in range(10)]
    print("\ncfg3 pose error vs GT per iteration:", ["%.2e" % e for e in errs_p])
    assert errs_p[-1] < errs_p[0]
    worst = (0.0, 0.0)
    for k in range(10):
        ep, ed = rel_err(P[k], z["poses"][k]), rel_err(D[k], z["disps"][k].astype(np.float64))
        worst = (max(worst[0], ep), max(worst[1], ed))
        assert ep < TOL and ed < TOL, f"iteration {k + 1}: poses {ep} disps {ed}"
    print(f"cfg3 iterations 1..10 vs fp64 sparse oracle: worst poses {worst[0]:.2e} disps {worst[1]:.2e}")


def test_davis_like_window_against_dense_oracle():
    """cfg2 stand-in (configs/davis_demo.yaml shape: 400 patches / frame, S_slam 12, kf_stride 2, 15-pose
    optimisation window): the reference's own graph bookkeeping replayed on synthetic tracks, the update()
    pairing (pose call on weights_pose, structure-only call on weights) x 2, against the fp64 oracle."""
    import synth
    from gpu_util import run_ours
    ps, w_all = synth.make_slam_problem(n_frames=25, patches_per_frame=400, seed=7, buffer_size=64)
    ws, so = [ps.weights, w_all] * 2, [False, True] * 2
    P, D = run_ours(ps, ws, so)
    P64, D64 = _oracle().run_sequence(ps, ws, so, torch.float64, mode="sparse")
    ep, ed = rel_err(P, P64), rel_err(D, D64)
    print(f"\ndavis-like ({ps.E} edges, fixedp {ps.fixedp}): poses {ep:.2e} disps {ed:.2e}")
    assert ep < TOL and ed < TOL


def test_1024_keyframe_graph_against_sparse_oracle():
    """cfg5 (1024 KF / 262 144 tracks / 4 980 736 edges, BASELINE.json configs[4]) on one device: banded reduced system
    with 6138 unknowns; poses and disparities of two chained iterations against the sparse fp64 oracle
    (tests/golden/cfg5_x2_sparse64.npz, produced by tests/golden/make_golden_large.py — the dense reference cannot hold
    this graph: 6.4 GB per E-sized temporary)."""
    import synth
    from gpu_util import run_ours
    prob = synth.make_config("cfg5")
    z = _golden_large("cfg5_x2_sparse64.npz")
    assert int(z["edges"]) == prob.E
    P, D, plan, t = run_ours(prob, [prob.weights] * 2, [False] * 2, return_plan=True)
    assert np.isfinite(P).all() and np.isfinite(D).all()
    assert plan.info.banded == 1 and plan.info.n_total == 1024 and plan.info.block_bandwidth == 18
    assert plan.status() == 0
    for k in range(2):
        ep, ed = rel_err(P[k], z["poses"][k]), rel_err(D[k], z["disps"][k].astype(np.float64))
        print(f"\ncfg5 iteration {k + 1} vs fp64 sparse oracle: poses {ep:.2e} disps {ed:.2e}")
        assert ep < TOL and ed < TOL


def test_sintel_like_full_sequence_window():
    """cfg4 stand-in (configs/sintel.yaml: 256 patches / frame, S_slam 12, kf_stride 2; "full-sequence" = the removal and
    optimisation windows cover all 50 frames, so 48 free poses and every keyframe step's 18 432 edges stay in the
    graph): the reference's own bookkeeping replayed on synthetic tracks, update() pairing x 2, against the fp64 oracle."""
    import synth
    from gpu_util import run_ours
    ps, w_all = synth.make_slam_problem(n_frames=50, patches_per_frame=256, seed=4, buffer_size=64, opt_window=64,
                                        removal_window=64, width=1024, height=436, name="sintel_like")
    assert ps.fixedp == 1 and ps.E > 350000
    ws, so = [ps.weights, w_all] * 2, [False, True] * 2
    P, D = run_ours(ps, ws, so)
    P64, D64 = _oracle().run_sequence(ps, ws, so, torch.float64, mode="sparse")
    ep, ed = rel_err(P, P64), rel_err(D, D64)
    print(f"\nsintel-like ({ps.E} edges, 48 free poses): poses {ep:.2e} disps {ed:.2e}")
    assert ep < TOL and ed < TOL


def test_structure_only_and_ba_variant_on_mid_graph():
    import synth
    from gpu_util import run_ours
    prob = synth.make_config("mid")
    ws, so = [prob.weights] * 3, [True, False, True]
    for variant in ("rgbd", "ba"):
        P, D = run_ours(prob, ws, so, variant=variant)
        P64, D64 = _oracle().run_sequence(prob, ws, so, torch.float64, variant=variant, mode="sparse")
        ep, ed = rel_err(P, P64), rel_err(D, D64)
        print(f"\nmid {variant} [so, full, so]: poses {ep:.2e} disps {ed:.2e}")
        assert ep < TOL and ed < TOL


def test_fused_update_equals_call_sequence():
    """BA_update (one native call for the loop of main/batrack.py:869-875) against the same steps issued one by one
    and against the fp64 oracle."""
    import synth
    from batrack_b200.ba import BA_update
    from batrack_b200.lietorch import SE3
    from gpu_util import as_cuda, run_ours
    ps, w_all = synth.make_slam_problem(n_frames=21, patches_per_frame=64, seed=11)
    t = as_cuda(ps)
    wp = torch.from_numpy(ps.weights).cuda()[None]
    wa = torch.from_numpy(w_all).cuda()[None]
    G, p = BA_update(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], wp, wa,
                     ps.lmbda, t["ii"], t["jj"], t["kk"], ps.bounds, ep=ps.ep, fixedp=ps.fixedp, loss=ps.loss,
                     alpha=ps.alpha, iters=3)
    ws, so = [ps.weights, w_all] * 3, [False, True] * 3
    P, D = run_ours(ps, ws, so)
    P64, D64 = _oracle().run_sequence(ps, ws, so, torch.float64, mode="sparse")
    assert rel_err(G.data[0].cpu().numpy(), P[-1]) < 2e-6 and rel_err(p[0, :, 2, 0, 0].cpu().numpy(), D[-1]) < 2e-5
    assert rel_err(G.data[0].cpu().numpy(), P64[-1]) < TOL and rel_err(p[0, :, 2, 0, 0].cpu().numpy(), D64[-1]) < TOL


def test_point_cloud_and_flow_mag_match_oracle():
    """The reprojection consumers next to BA (main/batrack.py:891-893, 1011-1018)."""
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
    f = lambda a: torch.from_numpy(a).double()
    P, X, K = f(ps.poses), f(ps.patches), f(ps.intrinsics)
    ixc = ix.cpu()
    x0 = torch.stack([(X[:, 0] - K[ixc, 2]) / K[ixc, 0], (X[:, 1] - K[ixc, 3]) / K[ixc, 1], torch.ones(NM, dtype=torch.float64), X[:, 2]], 1)
    ref = se3_ops.se3_act4(se3_ops.se3_inv(P[ixc].contiguous()), x0)
    assert rel_err(pc[0, :, 0, 0].cpu().numpy(), ref.numpy()) < 2e-6
    fm = pops.flow_mag(SE3(t["poses"]), t["patches"], t["intrinsics"], t["ii"], t["jj"], t["kk"], beta=0.5)
    g = lambda a: torch.from_numpy(a).long()
    c0, *_ = _oracle().reproject_with_jacobians(P, X, K, g(ps.ii), g(ps.ii), g(ps.kk))
    c1, *_ = _oracle().reproject_with_jacobians(P, X, K, g(ps.ii), g(ps.jj), g(ps.kk))
    Pt = P.clone()                                                    # tonly: rotation of Gij dropped (projective_ops.py:63-64)
    Gij = se3_ops.se3_mul(P[g(ps.jj)].contiguous(), se3_ops.se3_inv(P[g(ps.ii)].contiguous()))
    Gij[:, 3:] = torch.tensor([0, 0, 0, 1.0])
    xk = X[g(ps.kk)]
    Ki, Kj = K[g(ps.ii)], K[g(ps.jj)]
    X0 = torch.stack([(xk[:, 0] - Ki[:, 2]) / Ki[:, 0], (xk[:, 1] - Ki[:, 3]) / Ki[:, 1], torch.ones(len(xk), dtype=torch.float64), xk[:, 2]], 1)
    X1 = se3_ops.se3_act4(Gij, X0)
    dcl = 1.0 / X1[:, 2].clamp(min=1e-2)
    c2 = torch.stack([Kj[:, 0] * dcl * X1[:, 0] + Kj[:, 2], Kj[:, 1] * dcl * X1[:, 1] + Kj[:, 3]], 1)
    ref_fm = 0.5 * (c1 - c0).norm(dim=1) + 0.5 * (c2 - c0).norm(dim=1)
    assert np.abs(fm[0, :, 0, 0].cpu().numpy() - ref_fm.numpy()).max() < 2e-3     # pixels; fp32 coordinates ~1e3


def test_host_buffer_pipeline_equals_device_call():
    """ba_step_host_async / ba_host_sync (pinned host arrays in and out, double-buffered staging): four pipelined
    steps with different weights give bitwise what BA_rgbd_droid gives on device tensors; ba_step_host likewise."""
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.host import HostBA
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import Plan
    from gpu_util import as_cuda
    ps, w_all = synth.make_slam_problem(n_frames=21, patches_per_frame=64, seed=5)
    t = as_cuda(ps)
    N, NM = t["poses"].shape[1], t["patches"].shape[1]
    plan = Plan(t["ii"], t["jj"], t["kk"], N, NM)
    host = {k: t[k].cpu().pin_memory() for k in ("poses", "patches", "patches_monodisp", "intrinsics", "targets_2d")}
    rng = np.random.default_rng(0)
    ws = [torch.from_numpy((ps.weights * rng.uniform(0.5, 1.0, ps.weights.shape)).astype(np.float32))[None].pin_memory()
          for _ in range(4)]
    outs = [(torch.empty(1, N, 7).pin_memory(), torch.empty(1, NM, 3, 1, 1).pin_memory()) for _ in range(4)]
    hba = HostBA(plan)
    for w, (op, oq) in zip(ws, outs):
        hba.submit(host["poses"], host["patches"], host["patches_monodisp"], host["intrinsics"], host["targets_2d"], w,
                   ps.lmbda, ps.bounds, op, oq, ep=ps.ep, fixedp=ps.fixedp, structure_only=False, loss=ps.loss,
                   alpha=ps.alpha)
    hba.sync()
    for w, (op, oq) in zip(ws, outs):
        G, p = BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                             w.cuda(), ps.lmbda, t["ii"], t["jj"], t["kk"], ps.bounds, ep=ps.ep, fixedp=ps.fixedp,
                             structure_only=False, loss=ps.loss, alpha=ps.alpha, plan=plan)
        # S / y are summed with fp64 atomics in a run-dependent order: equal to rounding, not bitwise
        assert rel_err(op.numpy(), G.data.cpu().numpy()) < 1e-6
        assert rel_err(oq.numpy(), p.cpu().numpy()) < 1e-6
    # dependent steps with the early upload of the step-independent inputs (ba_prefetch_host_async): step k+1 starts from
    # the host results of step k; the chain equals the same chain on device tensors
    chain = [(torch.empty(1, N, 7).pin_memory(), torch.empty(1, NM, 3, 1, 1).pin_memory()) for _ in range(3)]
    cur = (host["poses"], host["patches"])
    for k in range(3):
        hba.submit(cur[0], cur[1], host["patches_monodisp"], host["intrinsics"], host["targets_2d"], ws[k], ps.lmbda, ps.bounds,
                   chain[k][0], chain[k][1], ep=ps.ep, fixedp=ps.fixedp, structure_only=False, loss=ps.loss, alpha=ps.alpha)
        if k + 1 < 3:
            hba.prefetch(host["patches_monodisp"], host["intrinsics"], host["targets_2d"], ws[k + 1])
        hba.sync()
        cur = chain[k]
    G, p = SE3(t["poses"]), t["patches"]
    for k in range(3):
        G, p = BA_rgbd_droid(G, p, t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None, ws[k].cuda(), ps.lmbda, t["ii"], t["jj"],
                             t["kk"], ps.bounds, ep=ps.ep, fixedp=ps.fixedp, structure_only=False, loss=ps.loss, alpha=ps.alpha, plan=plan)
        assert rel_err(chain[k][0].numpy(), G.data.cpu().numpy()) < 1e-5 and rel_err(chain[k][1].numpy(), p.cpu().numpy()) < 1e-5
    # the same dependent chain enqueued back to back, no host wait in between: the library orders the upload of step k+1
    # after the download of step k into the same host arrays (host-buffer hazard tracking, ba_stage_host_async); two
    # output arrays used in turn, as a caller would
    ring = [(torch.empty(1, N, 7).pin_memory(), torch.empty(1, NM, 3, 1, 1).pin_memory()) for _ in range(2)]
    steps = 6
    cur = (host["poses"], host["patches"])
    for k in range(steps):
        hba.submit(cur[0], cur[1], host["patches_monodisp"], host["intrinsics"], host["targets_2d"], ws[k % 4], ps.lmbda, ps.bounds,
                   ring[k & 1][0], ring[k & 1][1], ep=ps.ep, fixedp=ps.fixedp, structure_only=False, loss=ps.loss, alpha=ps.alpha)
        if k + 1 < steps:
            hba.prefetch(host["patches_monodisp"], host["intrinsics"], host["targets_2d"], ws[(k + 1) % 4])
        cur = ring[k & 1]
    hba.sync()
    G, p = SE3(t["poses"]), t["patches"]
    for k in range(steps):
        G, p = BA_rgbd_droid(G, p, t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None, ws[k % 4].cuda(), ps.lmbda, t["ii"], t["jj"],
                             t["kk"], ps.bounds, ep=ps.ep, fixedp=ps.fixedp, structure_only=False, loss=ps.loss, alpha=ps.alpha, plan=plan)
    last = ring[(steps - 1) & 1]
    assert rel_err(last[0].numpy(), G.data.cpu().numpy()) < 1e-5 and rel_err(last[1].numpy(), p.cpu().numpy()) < 1e-5
    with pytest.raises(RuntimeError):
        hba.submit(t["poses"], host["patches"], host["patches_monodisp"], host["intrinsics"], host["targets_2d"], ws[0],
                   ps.lmbda, ps.bounds, *outs[0])


@pytest.mark.parametrize("stream", [0, 1])
@pytest.mark.parametrize("n_kf", [64, 112])
def test_band_solver_failure_is_silent_and_recoverable(n_kf, stream):
    """The band (DMMA) solver — plain at 64 keyframes, twisted over two CTAs at 112 (>= 64 tile columns) — launched
    behind the Schur kernel (default) or next to it (streaming hand-over, option "stream"): a failed factorisation (ba.py:9-13) leaves the poses unchanged
    while the depths still move, and the next call on the same plan is correct again (flags carry a new epoch)."""
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import Plan
    from gpu_util import as_cuda
    prob = synth.make_window_problem(n_kf, 64, 19, seed=3)
    t = as_cuda(prob)
    N, NM = t["poses"].shape[1], t["patches"].shape[1]
    plan = Plan(t["ii"], t["jj"], t["kk"], N, NM)
    assert plan.info.banded
    plan.set_option("stream", stream)
    w = torch.from_numpy(prob.weights).cuda()[None]

    def call(ep):
        return BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                             w, prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=ep, fixedp=prob.fixedp,
                             loss=prob.loss, alpha=prob.alpha, plan=plan)

    G0, p0 = call(prob.ep)
    assert plan.status() == 0
    Gf, pf = call(-1e12)                                                  # A = S - 1e12 I is not positive definite
    assert plan.status() & 1
    assert rel_err(Gf.data.cpu().numpy(), t["poses"].cpu().numpy()) < 1e-6
    assert torch.isfinite(pf).all() and not torch.equal(pf, t["patches"])
    G1, p1 = call(prob.ep)
    assert plan.status() == 0
    assert rel_err(G1.data.cpu().numpy(), G0.data.cpu().numpy()) < 1e-6 and rel_err(p1.cpu().numpy(), p0.cpu().numpy()) < 1e-6
    P64, D64 = _oracle().run_sequence(prob, [prob.weights], [False], torch.float64, mode="sparse")
    assert rel_err(G0.data[0].cpu().numpy(), P64[0]) < TOL and rel_err(p0[0, :, 2, 0, 0].cpu().numpy(), D64[0]) < TOL


@pytest.mark.parametrize("stream", [0, 1])
def test_ba_step_replays_from_a_cuda_graph(stream):
    """the step must be graph-capturable. A captured BA_rgbd_droid call (plain, and with the band solver
    streamed next to the Schur kernel: cross-stream events, no host-side per-call state) is replayed on new weights and
    matches eager calls."""
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import Plan
    from gpu_util import as_cuda
    prob = synth.make_window_problem(64, 64, 19, seed=5)
    t = as_cuda(prob)
    N, NM = t["poses"].shape[1], t["patches"].shape[1]
    plan = Plan(t["ii"], t["jj"], t["kk"], N, NM)
    plan.set_option("stream", stream)
    w = torch.from_numpy(prob.weights).cuda()[None].clone()

    def call():
        return BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                             w, prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=prob.ep, fixedp=prob.fixedp,
                             loss=prob.loss, alpha=prob.alpha, plan=plan)

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                                         # warm-up on the capture stream (streams / events / tables)
        for _ in range(2):
            call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        G, p = call()
    rng = np.random.default_rng(1)
    for _ in range(3):
        w.copy_(torch.from_numpy((prob.weights * rng.uniform(0.3, 1.0, prob.weights.shape)).astype(np.float32)).cuda()[None])
        g.replay()
        torch.cuda.synchronize()
        Gr, pr = G.data.clone(), p.clone()
        Ge, pe = call()
        assert rel_err(Gr.cpu().numpy(), Ge.data.cpu().numpy()) < 1e-6
        assert rel_err(pr.cpu().numpy(), pe.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("n_kf", [12, 40])
def test_every_band_solver_matches_the_oracle(n_kf):
    """The reduced solve has four implementations behind BA_OPT_SOLVER (shared-memory tile solver: short systems and
    bands up to 145; DMMA band solver: long bands up to 120; scalar window solver; automatic choice). On one graph they
    must all land on the fp64 oracle's result, and a failed factorisation (ba.py:9-13) must leave the poses alone."""
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import Plan
    from gpu_util import as_cuda
    prob = synth.make_window_problem(n_kf, 64, 19, seed=11)
    t = as_cuda(prob)
    N, NM = t["poses"].shape[1], t["patches"].shape[1]
    w = torch.from_numpy(prob.weights).cuda()[None]
    P64, D64 = _oracle().run_sequence(prob, [prob.weights], [False], torch.float64, mode="sparse")
    for sv in ("auto", "tiles", "diag", "window"):
        plan = Plan(t["ii"], t["jj"], t["kk"], N, NM)
        plan.set_option("solver", sv)

        def call(ep):
            return BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                                 w, prob.lmbda, t["ii"], t["jj"], t["kk"], prob.bounds, ep=ep, fixedp=prob.fixedp,
                                 loss=prob.loss, alpha=prob.alpha, plan=plan)

        G, p = call(prob.ep)
        assert plan.status() == 0, sv
        assert rel_err(G.data[0].cpu().numpy(), P64[0]) < TOL, sv
        assert rel_err(p[0, :, 2, 0, 0].cpu().numpy(), D64[0]) < TOL, sv
        Gf, pf = call(-1e12)
        assert plan.status() & 1, sv
        assert rel_err(Gf.data.cpu().numpy(), t["poses"].cpu().numpy()) < 1e-6, sv
        G2, _ = call(prob.ep)
        assert plan.status() == 0 and rel_err(G2.data.cpu().numpy(), G.data.cpu().numpy()) < 1e-6, sv


@pytest.mark.parametrize("cfg,world", [("cfg1", 2), ("mid", 4)])
def test_sharded_assemble_and_solve_through_the_c_abi(cfg, world):
    """SURVEY.md §8e on hardware without a second GPU: `world` keyframe-window shards of one graph, each with its own plan
    on the same device, go through the sharded entry points — ba_plan_set_layout (agreed pose count / band width),
    ba_assemble (partial [S | y]), the exchange (here a plain sum of the ranks' exchange buffers, what the NCCL
    all-reduce computes), ba_solve_update — and the merged result must match the fp64 oracle of the full graph."""
    import ctypes as C
    import synth
    from batrack_b200 import _capi
    from batrack_b200.ba import _problem
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import Plan
    from gpu_util import as_cuda
    n_kf = synth.CONFIGS[cfg][0]
    full = synth.make_config(cfg)
    shards = [synth.make_config(cfg, kf_lo=(n_kf * r) // world, kf_hi=(n_kf * (r + 1)) // world) for r in range(world)]
    ts = [as_cuda(sh) for sh in shards]
    N, NM = ts[0]["poses"].shape[1], ts[0]["patches"].shape[1]
    plans = [Plan(t["ii"], t["jj"], t["kk"], N, NM) for t in ts]
    n_total, bwb = max(p.info.n_total for p in plans), max(p.info.block_bandwidth for p in plans)
    for p in plans:
        p.set_layout(n_total, bwb)
    L, st = _capi.lib(), _capi.stream_ptr(torch.device("cuda:0"))
    probs = []
    for sh, t, p in zip(shards, ts, plans):
        w = torch.from_numpy(sh.weights).cuda()[None]
        probs.append(_problem(p, SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], w, sh.lmbda,
                              sh.bounds, sh.ep, sh.fixedp, False, sh.loss, sh.alpha))
        _capi.check(L.ba_assemble(p.handle, C.byref(probs[-1][0]), st), "ba_assemble")
    bufs = [p.reduced_system() for p in plans]
    assert len({b.numel() for b in bufs}) == 1                      # the same layout on every rank
    total = torch.stack(bufs).sum(0)
    for b in bufs:
        b.copy_(total)
    disp = ts[0]["patches"][0, :, 2, 0, 0].clamp(1e-3, 10.0).double()
    base = disp.clone()
    poses = None
    for sh, p, pr in zip(shards, plans, probs):
        _capi.check(L.ba_solve_update(p.handle, C.byref(pr[0]), st), "ba_solve_update")
        assert p.status() == 0
        disp += pr[2][0, :, 2, 0, 0].double() - base               # every rank moves its own tracks only
        poses = pr[1] if poses is None else poses
        assert torch.equal(poses, pr[1])                            # replicated solve: bitwise identical
    P64, D64 = _oracle().run_sequence(full, [full.weights], [False], torch.float64, mode="sparse")
    assert rel_err(poses[0].cpu().numpy(), P64[0]) < TOL
    assert rel_err(disp.cpu().numpy(), D64[0]) < TOL
