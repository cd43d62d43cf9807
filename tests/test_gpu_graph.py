"""SURVEY.md §8 f3 — factor-graph maintenance without host synchronisation: a capacity plan re-derived on the device
(ba_plan_update) must equal a plan built from scratch (ba_plan_create) for every graph of a SLAM session."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def slam_graph_sequence(n_frames, M, s_slam=12, kf_stride=2, removal_window=20):
    """The reference's bookkeeping replayed step by step (main/batrack.py:990 every kf_stride-th step appends the
    __edges() of :399-410 through append_factors :189-204; keyframe_simple :1023-1026 drops edges whose source frame
    left the removal window). Yields (n, ii, jj, kk) after every step that changed the graph."""
    ii = jj = kk = np.zeros(0, dtype=np.int64)
    for step in range(2, n_frames + 1):
        if (step - 1) % kf_stride != 0:
            continue
        lo = max(step - s_slam, 0)
        kf = np.arange(lo, step, kf_stride)
        pk = (kf[:, None] * M + np.arange(M)[None, :]).reshape(-1)
        fr = np.arange(lo, step)
        k_new = np.repeat(pk, fr.shape[0])
        j_new = np.tile(fr, pk.shape[0])
        ii = np.concatenate([ii, k_new // M]); jj = np.concatenate([jj, j_new]); kk = np.concatenate([kk, k_new])
        keep = ii >= step - removal_window
        ii, jj, kk = ii[keep], jj[keep], kk[keep]
        yield step, ii.copy(), jj.copy(), kk.copy()


def _random_problem(rng, N, NM, ii, jj, kk, dev):
    """Inputs of one BA call on the graph (geometry does not matter for a plan-vs-plan comparison)."""
    import synth
    gt = synth._exp(np.cumsum(rng.normal(size=(N, 6)) * np.array([0.02] * 3 + [0.004] * 3), axis=0))
    intr = np.tile(np.array([800.0, 800.0, 480.0, 270.0]), (N, 1))
    xyd = np.stack([rng.uniform(30, 930, NM), rng.uniform(30, 510, NM), rng.uniform(0.2, 1.2, NM)], 1)
    tgt, _ = synth._reproject(gt, xyd, intr, ii, jj, kk)
    tgt = tgt + rng.normal(size=tgt.shape)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=torch.float32)
    return dict(poses=f(gt)[None], patches=f(xyd).view(1, NM, 3, 1, 1), mono=f(xyd[:, 2]).view(1, NM, 1), intr=f(intr)[None],
                targets=f(tgt)[None], weights=torch.ones(1, ii.shape[0], 2, device=dev))


def test_capacity_plan_update_equals_plan_from_scratch():
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    from batrack_b200.plan import CapacityPlan, Plan
    dev = torch.device("cuda:0")
    n_frames, M, N = 25, 96, 32
    NM = N * M
    seq = list(slam_graph_sequence(n_frames, M))
    cap_E = max(g[1].shape[0] for g in seq)
    cplan = CapacityPlan(N, NM, cap_edges=cap_E, cap_tracks=NM, cap_groups=N, cap_pattern=N * 128)
    rng = np.random.default_rng(0)
    fields = ("n_edges", "n_total", "n_tracks", "n_groups", "n_chunks", "max_degree", "max_slots", "block_bandwidth",
              "perm_identity", "banded")
    for step, ii, jj, kk in seq:
        t = [torch.from_numpy(a).to(dev) for a in (ii, jj, kk)]
        ref = Plan(*t, N, NM)
        cplan.update(*t).finalize()
        for f in fields:
            assert getattr(cplan.info, f) == getattr(ref.info, f), (step, f, getattr(cplan.info, f), getattr(ref.info, f))
        assert torch.equal(cplan.tracks(), ref.tracks())                       # == torch.unique(kk), bit-exact
        assert torch.equal(cplan.tracks().long(), torch.unique(t[2]))
        d = _random_problem(rng, N, NM, ii, jj, kk, dev)
        fixedp = max(step - 15, 1)
        outs = []
        for plan in (ref, cplan):
            G, p = BA_rgbd_droid(SE3(d["poses"]), d["patches"], d["mono"], d["intr"], d["targets"], None, d["weights"], 1e-4,
                                 *t, [0, 0, 960, 540], ep=10.0, fixedp=fixedp, loss="huber", alpha=0.05, plan=plan)
            assert plan.status() == 0
            outs.append((G.data.clone(), p.clone()))
        assert (outs[0][0] - outs[1][0]).abs().max() <= 1e-6 * outs[0][0].abs().max()
        assert (outs[0][1] - outs[1][1]).abs().max() <= 1e-6 * outs[0][1].abs().max()


def test_capacity_plan_reports_overflow_and_recovers():
    from batrack_b200.plan import CapacityPlan
    dev = torch.device("cuda:0")
    M, N = 32, 16
    seq = list(slam_graph_sequence(13, M))
    _, ii, jj, kk = seq[-1]
    t = [torch.from_numpy(a).to(dev) for a in (ii, jj, kk)]
    small = CapacityPlan(N, N * M, cap_edges=ii.shape[0], cap_tracks=N * M, cap_groups=2, cap_pattern=64)
    with pytest.raises(RuntimeError, match="capacity"):
        small.update(*t).finalize()
    _, i0, j0, k0 = seq[0]                                                    # a graph that fits: the plan is usable again
    t0 = [torch.from_numpy(a).to(dev) for a in (i0, j0, k0)]
    small.update(*t0).finalize()
    assert small.info.n_edges == i0.shape[0] and torch.equal(small.tracks().long(), torch.unique(t0[2]))


def test_plan_update_uses_the_device_side_edge_count():
    """n_edges_dev: the live edge count stays on the device (a graph compacted by a device-side removal); entries of the
    index arrays beyond it are ignored."""
    from batrack_b200.plan import CapacityPlan, Plan
    dev = torch.device("cuda:0")
    M, N = 48, 24
    _, ii, jj, kk = list(slam_graph_sequence(21, M))[-1]
    E = ii.shape[0]
    live = E - 777
    pad = lambda a: torch.from_numpy(np.concatenate([a[:live], np.full(777, 10 ** 6, np.int64)])).to(dev)
    cplan = CapacityPlan(N, N * M, cap_edges=E, cap_groups=N, cap_pattern=N * 128)
    cplan.update(pad(ii), pad(jj), pad(kk), n_edges_dev=torch.tensor([live], dtype=torch.int32, device=dev)).finalize()
    ref = Plan(*[torch.from_numpy(a[:live]).to(dev) for a in (ii, jj, kk)], N, N * M)
    assert cplan.info.n_edges == live == ref.info.n_edges
    assert cplan.info.n_groups == ref.info.n_groups and torch.equal(cplan.tracks(), ref.tracks())


def test_device_factor_graph_replays_the_reference_bookkeeping():
    """FactorGraph (ba_graph_*): append_factors / removal window / keyframe() edge surgery / arbitrary masks on the device
    against the same operations in numpy (the reference's torch.cat / boolean-mask code, main/batrack.py:189-212,
    1042-1051, 1072-1073), payload rows included; the plan derived from the device-resident graph equals a plan built from
    scratch on the host's copy."""
    from batrack_b200.graph import FactorGraph
    from batrack_b200.plan import Plan
    dev = torch.device("cuda:0")
    n_frames, M, N, s_slam, stride, removal = 23, 64, 32, 12, 2, 14
    rng = np.random.default_rng(3)
    # a random mask breaks the "one pattern group per keyframe" structure: room for one group per track
    fg = FactorGraph(N, M, cap_edges=200000, cap_groups=N * M, cap_pattern=200000, device=dev)
    ix = torch.arange(N * M, device=dev) // M                                   # patch -> source frame (self.ix)
    ii = jj = kk = np.zeros(0, dtype=np.int64)
    tg, w, wp = np.zeros((0, 3), np.float32), np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32)
    n = 0

    def check(tag):
        gi, gj, gk, gt, gw, gwp = fg.edges()
        assert gi.numel() == ii.shape[0], (tag, gi.numel(), ii.shape[0])
        for a, b in ((gi, ii), (gj, jj), (gk, kk), (gt, tg), (gw, w), (gwp, wp)):
            assert np.array_equal(a.cpu().numpy(), b), tag

    for step in range(2, n_frames + 1):
        n = step
        if (step - 1) % stride == 0:
            lo = max(step - s_slam, 0)
            kf = np.arange(lo, step, stride)
            pk = (kf[:, None] * M + np.arange(M)[None, :]).reshape(-1)
            fr = np.arange(lo, step)
            k_new, j_new = np.repeat(pk, fr.shape[0]), np.tile(fr, pk.shape[0])
            t_new = rng.normal(size=(k_new.shape[0], 3)).astype(np.float32)
            w_new = rng.uniform(size=(k_new.shape[0], 2)).astype(np.float32)
            wp_new = rng.uniform(size=(k_new.shape[0], 2)).astype(np.float32)
            c = lambda a: torch.from_numpy(a).to(dev)
            fg.append_factors(c(k_new), c(j_new), ix, c(t_new), c(w_new), c(wp_new))
            ii = np.concatenate([ii, k_new // M]); jj = np.concatenate([jj, j_new]); kk = np.concatenate([kk, k_new])
            tg, w, wp = np.concatenate([tg, t_new]), np.concatenate([w, w_new]), np.concatenate([wp, wp_new])
        if step == 15:                                                          # keyframe(): frame k leaves the window
            k = step - 4
            keep = ~((ii == k) | (jj == k))
            ii, jj, kk, tg, w, wp = ii[keep], jj[keep], kk[keep], tg[keep], w[keep], wp[keep]
            kk = np.where(ii > k, kk - M, kk); ii = np.where(ii > k, ii - 1, ii); jj = np.where(jj > k, jj - 1, jj)
            fg.remove_keyframe(k)
        if step == 18:                                                          # an arbitrary mask
            m = rng.uniform(size=fg.n_upper) < 0.1
            live = m[:ii.shape[0]]
            ii, jj, kk, tg, w, wp = ii[~live], jj[~live], kk[~live], tg[~live], w[~live], wp[~live]
            fg.remove_factors(torch.from_numpy(m).to(dev))
        keep = ii >= n - removal                                                # removal window
        ii, jj, kk, tg, w, wp = ii[keep], jj[keep], kk[keep], tg[keep], w[keep], wp[keep]
        fg.remove_before(n - removal, ix)
        if step % 3 == 0 or step == n_frames:
            plan = fg.plan()                                                    # enqueued; finalized by edges()
            check(step)
            ref = Plan(*[torch.from_numpy(a).to(dev) for a in (ii, jj, kk)], N, N * M)
            for f in ("n_edges", "n_total", "n_tracks", "n_groups", "n_chunks", "max_degree", "max_slots", "block_bandwidth"):
                assert getattr(plan.info, f) == getattr(ref.info, f), (step, f)
            assert torch.equal(plan.tracks(), ref.tracks())
    assert ii.shape[0] > 10000
