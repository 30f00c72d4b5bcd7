"""The host layer's own PNG / JPEG decoders (host/ckd_image.cpp; SURVEY 8 row f3, image.cpp:31-73 without DevIL) against
committed fixtures whose expected pixels were read back with Pillow (tests/golden/make_image_fixtures.py), and -- where the
reference tree is present -- against every image the reference loads (refdata/assets.npz, decoded by Pillow once)."""
import ctypes as C
import os

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMAGES = os.path.join(REPO, "tests", "golden", "images")


@pytest.fixture(scope="module")
def lib():
    from cookiedough_b200 import capi
    L = C.CDLL(capi.LIB_PATH)
    L.ckdhost_image_load.restype = C.c_void_p
    L.ckdhost_image_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ckdhost_image_free.argtypes = [C.c_void_p]
    L.ckdhost_set_asset_root.argtypes = [C.c_char_p]
    L.ckdhost_last_error.restype = C.c_char_p
    return L


def decode(L, path, gray):
    w, h = C.c_int(), C.c_int()
    p = L.ckdhost_image_load(path.encode(), int(gray), C.byref(w), C.byref(h))
    if not p:
        return None
    n = w.value * h.value
    try:
        buf = (C.c_uint8 * (n * (1 if gray else 4))).from_address(p)
        return np.frombuffer(buf, dtype=np.uint8 if gray else np.uint32).reshape(h.value, w.value).copy()
    finally:
        L.ckdhost_image_free(p)


EXPECTED = np.load(os.path.join(IMAGES, "expected.npz"))
FIXTURES = sorted({k.split(":")[0] for k in EXPECTED.files})


@pytest.mark.parametrize("name", FIXTURES)
def test_decoder_matches_fixture(lib, name):
    lib.ckdhost_set_asset_root(IMAGES.encode())
    bgra = decode(lib, name, False)
    assert bgra is not None, lib.ckdhost_last_error().decode()
    assert bgra.shape == EXPECTED[name + ":bgra"].shape
    assert np.array_equal(bgra, EXPECTED[name + ":bgra"]), f"{name}: {np.count_nonzero(bgra != EXPECTED[name + ':bgra'])} BGRA pixels differ"
    l8 = decode(lib, name, True)
    assert l8 is not None and np.array_equal(l8, EXPECTED[name + ":l8"]), f"{name}: luminance differs"


def test_fixture_set_covers_the_formats():
    names = set(FIXTURES)
    assert {"rgb8.png", "rgba8.png", "grey8.png", "greyalpha8.png", "grey1.png", "grey16.png", "rgb16.png", "palette8_trns.png", "palette4.png",
            "palette2.png", "rgb8_colourkey.png", "rgba8_adam7.png", "rgba16_adam7.png", "base_444.jpg", "base_422.jpg", "base_420.jpg",
            "prog_444.jpg", "prog_420.jpg", "base_420_restart.jpg", "prog_444_restart.jpg", "grey.jpg", "grey_prog.jpg", "narrow_420.jpg"} <= names


def test_errors_are_reported_like_the_reference(lib, tmp_path):
    lib.ckdhost_set_asset_root(str(tmp_path).encode())
    assert decode(lib, "missing.png", False) is None
    assert lib.ckdhost_last_error().decode().startswith("Can not load image: missing.png")      # image.cpp:40
    good = open(os.path.join(IMAGES, "rgb8.png"), "rb").read()
    (tmp_path / "crc.png").write_bytes(good[:60] + bytes([good[60] ^ 0x40]) + good[61:])
    assert decode(lib, "crc.png", False) is None and "CRC" in lib.ckdhost_last_error().decode()
    (tmp_path / "short.png").write_bytes(good[:len(good) // 2])
    assert decode(lib, "short.png", False) is None
    jpeg = open(os.path.join(IMAGES, "base_444.jpg"), "rb").read()
    (tmp_path / "short.jpg").write_bytes(jpeg[:200])
    assert decode(lib, "short.jpg", False) is None
    (tmp_path / "text.png").write_bytes(b"not an image at all, just bytes" * 4)
    assert decode(lib, "text.png", False) is None and "not a PNG" in lib.ckdhost_last_error().decode()


def test_every_reference_asset_decodes_to_the_shared_pixels(lib):
    """all 105 files the reference loads (90 PNG, 12 baseline + 3 progressive JPEG): the native decoders return exactly the
    pixels both sides of the parity tests share (Pillow's decode, oracle/build_ref.py prepare_assets)"""
    target = os.path.join(os.environ.get("CKD_REFERENCE", "/root/reference"), "target")
    npz_path = os.path.join(REPO, "refdata", "assets.npz")
    if not os.path.isdir(os.path.join(target, "assets")) or not os.path.exists(npz_path):
        pytest.skip("reference tree not present on this machine")
    from cookiedough_b200.assets import SPEC
    npz = np.load(npz_path)
    lib.ckdhost_set_asset_root(target.encode())
    checked = 0
    for path, spec in SPEC.items():
        if path not in npz.files:
            continue
        got = decode(lib, path, bool(spec[2]))
        assert got is not None, lib.ckdhost_last_error().decode()
        assert got.shape == npz[path].shape and np.array_equal(got, npz[path]), path
        checked += 1
    assert checked >= 100


def test_mutated_files_decode_or_fail_cleanly(lib, tmp_path):
    """seeded mutations of every fixture (the ASAN/UBSAN build of the same loop is tests/tools/fuzz_image_decode.py:
    12,000 mutants clean): the decoder either returns pixels or reports an error, it never takes the process down"""
    import random
    import sys
    sys.path.insert(0, os.path.join(REPO, "tests", "tools"))
    from fuzz_image_decode import mutate
    rng = random.Random(7)
    lib.ckdhost_set_asset_root(str(tmp_path).encode())
    decoded = failed = 0
    for name in FIXTURES:
        data = open(os.path.join(IMAGES, name), "rb").read()
        for m in range(12):
            (tmp_path / "m.bin").write_bytes(mutate(data, rng, name.endswith(".png")))
            for gray in (False, True):
                if decode(lib, "m.bin", gray) is None:
                    failed += 1
                    assert lib.ckdhost_last_error().decode().startswith("Can not load image")
                else:
                    decoded += 1
    assert decoded > 50 and failed > 50


def decode_for(L, path, gray, res_x, res_y):
    L.ckdhost_image_load_for_resolution.restype = C.c_void_p
    L.ckdhost_image_load_for_resolution.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    w, h = C.c_int(), C.c_int()
    p = L.ckdhost_image_load_for_resolution(path.encode(), int(gray), res_x, res_y, C.byref(w), C.byref(h))
    if not p:
        return None
    try:
        buf = (C.c_uint8 * (w.value * h.value * (1 if gray else 4))).from_address(p)
        return np.frombuffer(buf, dtype=np.uint8 if gray else np.uint32).reshape(h.value, w.value).copy()
    finally:
        L.ckdhost_image_free(p)


def test_resolution_rules_on_files(lib, tmp_path):
    """the 4K rules of the host layer against their Python twin (cookiedough_b200/assets.py): output-sized art,
    FX-map sized maps and the ribbon strip are nearest-resampled, everything else keeps its size, and the tunnelscape colour
    map the reference's checkout lacks is the landscape's at twice the size"""
    from PIL import Image
    from cookiedough_b200.assets import _nearest_resize

    def write(path, arr):
        out = tmp_path / path
        out.parent.mkdir(parents=True, exist_ok=True)
        bgra = arr.view(np.uint8).reshape(arr.shape[0], arr.shape[1], 4)
        Image.fromarray(np.ascontiguousarray(bgra[..., [2, 1, 0, 3]]), "RGBA").save(out, "PNG", compress_level=1)

    rng = np.random.default_rng(5)
    files = {"assets/x/layer.png": (720, 1280), "assets/x/blurmap.png": (364, 644), "assets/demo/ribbons.png": (720, 2160),
             "assets/x/credits.png": (568, 1280), "assets/x/sprite.png": (128, 128), "assets/scape/C17W-edit.png": (64, 64)}
    arrays = {p: rng.integers(0, 2**32, size=hw, dtype=np.uint64).astype(np.uint32) for p, hw in files.items()}
    for p, a in arrays.items():
        write(p, a)
    lib.ckdhost_set_asset_root(str(tmp_path).encode())
    for res_x, res_y in ((1280, 720), (1920, 1080), (3840, 2160), (2560, 1080)):
        fx_x, fx_y = res_x // 2 + 4, res_y // 2 + 4
        want = {"assets/x/layer.png": _nearest_resize(arrays["assets/x/layer.png"], res_y, res_x),
                "assets/x/blurmap.png": _nearest_resize(arrays["assets/x/blurmap.png"], fx_y, fx_x),
                "assets/demo/ribbons.png": _nearest_resize(arrays["assets/demo/ribbons.png"], 720 * res_y // 720, 2160 * res_y // 720),
                "assets/x/credits.png": arrays["assets/x/credits.png"], "assets/x/sprite.png": arrays["assets/x/sprite.png"],
                "assets/scape/tscape-C7W-edit.png": _nearest_resize(arrays["assets/scape/C17W-edit.png"], 128, 128)}
        for p, ref in want.items():
            got = decode_for(lib, p, False, res_x, res_y)
            assert got is not None, lib.ckdhost_last_error().decode()
            assert got.shape == ref.shape and np.array_equal(got, ref), f"{p} at {res_x}x{res_y}"
    assert decode_for(lib, "assets/x/absent.png", False, 3840, 2160) is None


def test_every_reference_asset_at_4k_matches_the_harness(lib):
    target = os.path.join(os.environ.get("CKD_REFERENCE", "/root/reference"), "target")
    npz_path = os.path.join(REPO, "refdata", "assets.npz")
    if not os.path.isdir(os.path.join(target, "assets")) or not os.path.exists(npz_path):
        pytest.skip("reference tree not present on this machine")
    from cookiedough_b200.assets import SPEC, Assets
    lib.ckdhost_set_asset_root(target.encode())
    assets = Assets(3840, 2160)
    for path, spec in SPEC.items():
        got = decode_for(lib, path, bool(spec[2]), 3840, 2160)
        assert got is not None, lib.ckdhost_last_error().decode()
        want = assets[path]
        assert got.shape == want.shape and np.array_equal(got, want), path
        assets.drop(path)
