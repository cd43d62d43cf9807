import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Fixture:
    """A golden .npz (tests/golden/make_golden.py) exposed with the BAProblem field names."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        for k in z.files:
            setattr(self, k, z[k])
        if "scalars" in z.files:
            self.fixedp = int(z["scalars"][0])
            self.ep, self.lmbda, self.alpha = (float(v) for v in z["scalars"][1:])
            self.bounds = [float(v) for v in z["bounds"]]
            self.loss = "huber"


# fixture name -> (variant, loss override, uses lmbda_vec)
BA_FIXTURES = {
    "cfg1_rgbd": ("rgbd", None),
    "cfg1_ba": ("ba", None),
    "cfg1_cauchy": ("rgbd", "cauchy"),
    "cfg1_trivial_so": ("rgbd", "trivial"),
    "cfg1_lmbda_tensor": ("rgbd", None),
    "slam_dual": ("rgbd", None),
    "random_rgbd": ("rgbd", None),
    "random2_ba": ("ba", None),
    "tiny_all_fixed": ("rgbd", None),
    "tiny_bounds": ("rgbd", None),
}


@pytest.fixture
def golden():
    return Fixture


def rel_err(a, b):
    """max-abs-diff / max-abs (the yardstick BASELINE.md §5 uses)."""
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))
