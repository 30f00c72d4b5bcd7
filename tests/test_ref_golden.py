"""The oracle is pinned: the compiled reference (oracle/_ref) must reproduce the committed golden fixtures.
Runs the generator's child modes in separate processes (the reference keeps global state) and diffs the results."""
import json
import os
import subprocess
import sys
import tempfile

import pytest

from conftest import GOLDEN, REPO

from oracle import ref as oref

pytestmark = pytest.mark.skipif(not oref.available(720), reason="oracle/_ref not built (python oracle/build_ref.py)")

SCRIPT = os.path.join(GOLDEN, "make_golden.py")


def _regen(mode):
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, mode + ".json")
        subprocess.check_call([sys.executable, SCRIPT, "--child", mode, out], stdout=subprocess.DEVNULL, cwd=REPO)
        with open(out) as f:
            return json.load(f)


def _same_cpu(golden_rsqrt):
    """float goldens depend on the CPU's RSQRTPS table: only comparable on a CPU with the same table"""
    import numpy as np
    from cookiedough_b200.assets import Assets
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle import ref as o; from cookiedough_b200.assets import Assets;"
            "R = o.Reference(720, Assets(1280, 720, force_synthetic=True)); np.save(sys.argv[1], R.rsqrt_table(stride=1 << 13))" % REPO)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "t.npy")
        subprocess.check_call([sys.executable, "-c", code, p], cwd=REPO)
        return bool(np.array_equal(np.load(p), golden_rsqrt))


@pytest.mark.parametrize("mode", ["timeline", "scenario"])
def test_reference_reproduces_effect_goldens(mode, golden_effects, golden_rsqrt):
    from util import INTEGER_EFFECTS
    same_cpu = _same_cpu(golden_rsqrt)
    fresh = _regen(mode)
    pinned = golden_effects[mode]
    assert set(fresh) == set(pinned)
    for label, case in pinned.items():
        assert fresh[label]["params"] == pytest.approx(case["params"]), label
        if case["effect"] in INTEGER_EFFECTS or same_cpu:
            assert fresh[label]["sha256"] == case["sha256"], label


def test_reference_reproduces_post_goldens(golden_post):
    fresh = _regen("post")
    assert fresh == golden_post
