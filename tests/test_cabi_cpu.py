"""CPU (no GPU): the C-ABI library builds, loads, and exports every symbol include/batrack_ba.h declares;
host-side contract checks that need no device."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "batrack_ba.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|int64_t|char)\s*\*?\s*(\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from batrack_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25, declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/batrack_ba.h but not exported"
    assert sorted(_capi.SYMBOLS) == declared, "ctypes table and header disagree"
    assert _capi.lib().ba_version() == 100
    assert _capi.lib().ba_error_string(-3).decode() == "edge index out of range"


def test_struct_layout_matches_header():
    """BaProblem / BaPlanInfo are mirrored field by field in _capi.py; sizes must match the C compiler's."""
    import subprocess
    import tempfile
    from batrack_b200 import _capi
    code = '#include <stdio.h>\n#include "batrack_ba.h"\nint main(){printf("%zu %zu\\n", sizeof(BaProblem), sizeof(BaPlanInfo));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(code)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        a, b = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    assert int(a) == ctypes.sizeof(_capi.BaProblem)
    assert int(b) == ctypes.sizeof(_capi.BaPlanInfo)


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: the operator raises on host tensors instead of silently computing there."""
    import synth
    from batrack_b200.ba import BA_rgbd_droid
    from batrack_b200.lietorch import SE3
    prob = synth.make_config("tiny")
    t = prob.as_torch()
    with pytest.raises(RuntimeError, match="CUDA"):
        BA_rgbd_droid(SE3(t["poses"]), t["patches"], t["patches_monodisp"], t["intrinsics"], t["targets_2d"], None,
                      t["weights"], 1e-4, t["ii"], t["jj"], t["kk"], prob.bounds)
    with pytest.raises(RuntimeError, match="CUDA"):
        SE3(t["poses"]).inv()


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under batrack_b200/ may reference it."""
    pkg = os.path.join(ROOT, "batrack_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)


def test_synth_shards_partition_the_graph():
    import synth
    import numpy as np
    full = synth.make_config("cfg1")
    parts = [synth.make_config("cfg1", kf_lo=lo, kf_hi=hi) for lo, hi in ((0, 3), (3, 8))]
    assert sum(p.E for p in parts) == full.E
    cat = lambda k: np.concatenate([getattr(p, k) for p in parts])
    order = np.lexsort((cat("jj"), cat("kk")))
    forder = np.lexsort((full.jj, full.kk))
    for k in ("ii", "jj", "kk"):
        assert np.array_equal(cat(k)[order], getattr(full, k)[forder])
    assert np.array_equal(cat("targets")[order], full.targets[forder])
    assert np.array_equal(parts[0].poses, full.poses) and np.array_equal(parts[1].patches, full.patches)


def test_verification_build_exports_the_same_abi():
    """tools/racecheck_verify.sh builds the band solver with -DBA_VERIFY_SYNC into libbatrack_ba_verify.so (the form
    compute-sanitizer's racecheck can follow; selected with BATRACK_B200_LIB). When it is there it must export every
    symbol the product library does."""
    import ctypes
    import os
    import pytest
    from batrack_b200 import _capi
    path = os.path.join(os.path.dirname(_capi.__file__), "libbatrack_ba_verify.so")
    if not os.path.exists(path):
        pytest.skip("verification build not present (bash tools/racecheck_verify.sh build)")
    lib = ctypes.CDLL(path)
    for name in _capi.SYMBOLS:
        assert hasattr(lib, name), name
