import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

# the reference keeps its state in globals: one instance per resolution and process.  The 720p one is created with the
# compositor (Demo_Create) so that the Demo_Draw tests can share it with everything else.
os.environ.setdefault("CKD_REF_DEMO", "720")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `pytest -m gpu`")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return os.path.exists("/dev/nvidia0")


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_effects():
    with open(os.path.join(GOLDEN, "golden_effects_720.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_post():
    with open(os.path.join(GOLDEN, "golden_post_720.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def golden_rsqrt():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "rsqrt_table_golden.npy"))


@pytest.fixture(scope="session")
def synth_assets():
    from cookiedough_b200.assets import Assets
    return Assets(1280, 720, force_synthetic=True)


@pytest.fixture(scope="session")
def ctx_synth(synth_assets, golden_rsqrt):
    """720p CUDA context with the synthetic assets and the RSQRTPS table the goldens were generated with"""
    from cookiedough_b200 import capi
    ctx = capi.Context(1280, 720, 0, synth_assets)
    ctx.set_rsqrt_table(golden_rsqrt, 13)
    yield ctx
    ctx.close()


def _live(res_y):
    """(Reference, Context) pair sharing the same assets (the reference's art when refdata/assets.npz exists) and the
    host CPU's own RSQRTPS table, or None when the compiled reference is not available"""
    from oracle import ref as oref
    if not oref.available(res_y):
        return None
    from cookiedough_b200 import capi
    from cookiedough_b200.assets import Assets
    res_x = res_y * 16 // 9
    assets = Assets(res_x, res_y)
    R = oref.Reference.get(res_y, assets)
    ctx = capi.Context(res_x, res_y, 0, R.assets)
    return R, ctx


@pytest.fixture(scope="session")
def live720():
    pair = _live(720)
    if pair is None:
        pytest.skip("oracle/_ref not built")
    yield pair
    pair[1].close()


@pytest.fixture(scope="session")
def live2160():
    pair = _live(2160)
    if pair is None:
        pytest.skip("oracle/_ref not built")
    yield pair
    pair[1].close()
