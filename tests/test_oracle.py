"""CPU: the oracle restatement (oracle/ba_oracle.py, oracle/se3_ops.py) against the golden vectors
produced by the reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import BA_FIXTURES, Fixture, rel_err
from oracle import ba_oracle, se3_ops


@pytest.mark.parametrize("name", sorted(BA_FIXTURES))
@pytest.mark.parametrize("mode", ["dense", "sparse"])
def test_ba_oracle_fp64_matches_reference(name, mode):
    fx = Fixture(name)
    variant, loss = BA_FIXTURES[name]
    lm = fx.lmbda_vec if hasattr(fx, "lmbda_vec") else None
    P, D = ba_oracle.run_sequence(fx, fx.weights_seq, fx.structure_seq, torch.float64, variant=variant,
                                  lmbda=lm, loss=loss, mode=mode)
    assert rel_err(P, fx.ref64_poses) < 1e-9
    assert rel_err(D, fx.ref64_disps) < 1e-9


@pytest.mark.parametrize("name", ["cfg1_rgbd", "slam_dual", "random_rgbd"])
def test_ba_oracle_fp32_within_reference_noise(name):
    fx = Fixture(name)
    variant, loss = BA_FIXTURES[name]
    P, D = ba_oracle.run_sequence(fx, fx.weights_seq, fx.structure_seq, torch.float32, variant=variant, loss=loss)
    # the reference's own fp32 run is this far from its fp64 run; the restatement must be no worse than 3x
    ref_p = max(rel_err(fx.ref32_poses, fx.ref64_poses), 1e-6)
    ref_d = max(rel_err(fx.ref32_disps, fx.ref64_disps), 1e-6)
    assert rel_err(P, fx.ref64_poses) < 3 * ref_p
    assert rel_err(D, fx.ref64_disps) < 3 * ref_d


def test_transform_jacobians_match_reference():
    fx, tr = Fixture("cfg1_rgbd"), Fixture("cfg1_transform")
    f = lambda a: torch.from_numpy(a).double()
    g = lambda a: torch.from_numpy(a).long()
    coords, v, Ji, Jj, Jz = ba_oracle.reproject_with_jacobians(
        f(fx.poses), f(fx.patches), f(fx.intrinsics), g(fx.ii), g(fx.jj), g(fx.kk))
    assert rel_err(coords, tr.coords) < 1e-12
    assert np.array_equal(v.numpy(), tr.valid)
    assert rel_err(Ji, tr.Ji) < 1e-12
    assert rel_err(Jj, tr.Jj) < 1e-12
    assert rel_err(Jz, tr.Jz[..., 0]) < 1e-12


def test_se3_golden():
    z = Fixture("se3_ops")
    f = lambda a: torch.from_numpy(a)
    assert rel_err(se3_ops.se3_exp(f(z.a)), z.X) < 1e-14
    assert rel_err(se3_ops.se3_inv(f(z.Xr)), z.inv) < 1e-14
    assert rel_err(se3_ops.se3_mul(f(z.Xr), f(z.Y)), z.mul) < 1e-14
    assert rel_err(se3_ops.se3_act4(f(z.Xr), f(z.p4)), z.act4) < 1e-14
    assert rel_err(se3_ops.se3_adjT(f(z.Xr), f(z.c)), z.adjT) < 1e-14
    assert rel_err(se3_ops.se3_adj(f(z.Xr), f(z.c)), z.adj) < 1e-14
    assert rel_err(se3_ops.se3_log(f(z.X)), z.log) < 1e-14
    assert rel_err(se3_ops.se3_mul(se3_ops.se3_exp(f(z.a)), f(z.Y)), z.retr) < 1e-14


def test_se3_identities():
    """The four forward identities of the reference's lietorch/run_tests.py:16-52 (fp64, atol 1e-8)."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(256, 6, generator=g, dtype=torch.float64) * torch.tensor([1, 1, 1, .5, .5, .5])  # |phi| < pi
    b = torch.randn(256, 6, generator=g, dtype=torch.float64)
    X = se3_ops.se3_exp(a)
    assert torch.allclose(se3_ops.se3_log(X), a, atol=1e-8)                                   # :16-21
    I = se3_ops.se3_mul(X, se3_ops.se3_inv(X))                                               # :23-28
    assert torch.allclose(se3_ops.se3_log(I), torch.zeros_like(a), atol=1e-8)
    lhs = se3_ops.se3_mul(X, se3_ops.se3_exp(b))                                             # :30-41
    rhs = se3_ops.se3_mul(se3_ops.se3_exp(se3_ops.se3_adj(X, b)), X)
    assert torch.allclose(se3_ops.se3_matrix(lhs), se3_ops.se3_matrix(rhs), atol=1e-8)
    p = torch.randn(256, 3, generator=g, dtype=torch.float64)                                # :44-52
    p4 = torch.cat([p, torch.ones(256, 1, dtype=torch.float64)], 1)
    assert torch.allclose(se3_ops.se3_act3(X, p), (se3_ops.se3_matrix(X) @ p4[:, :, None])[:, :3, 0], atol=1e-8)
