"""Input art / maps for the effects, as raw arrays shared byte-for-byte by the CUDA path and the CPU oracle.

The reference decodes its art with DevIL (code/image.cpp:31-73: 32-bit BGRA or 8-bit luminance, upper-left
origin).  Decoding is outside the hot path, so the harness works on pre-decoded arrays:

* ``oracle/_ref/assets.npz`` (written by ``oracle/build_ref.py`` from the reference's ``target/assets``) when present,
* deterministic procedural stand-ins of the same shapes otherwise (so tests can run from a bare checkout).

Resolution rules for builds other than 1280x720: output-sized art is
nearest-upscaled, the 644x364 blur maps are nearest-resampled to the FX-map size, and the missing tunnelscape
colour map (``.MISSING_LARGE_BLOBS``) is the landscape colour map nearest-upscaled 2x to 2048x2048.
"""
import os

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# path -> (height, width, is_gray, is_output_sized, is_fxmap_sized)
SPEC = {
    "assets/shadertoy/nytrik-hextexture.png": (1024, 1024, False, False, False),
    "assets/shadertoy/nytrik-hextexture-fx.png": (1024, 1024, False, False, False),
    "assets/shadertoy/close-up-blur-map-1.png": (364, 644, False, False, True),
    "assets/shadertoy/close-up-blur-map-2.png": (364, 644, False, False, True),
    "assets/scape/D17.png": (1024, 1024, True, False, False),
    "assets/scape/C17W-edit.png": (1024, 1024, False, False, False),
    "assets/scape/foggradient.jpg": (1, 256, False, False, False),
    "assets/scape/tscape-D7-edit.png": (2048, 2048, True, False, False),
    "assets/scape/tscape-C7W-edit.png": (2048, 2048, False, False, False),
    "assets/ball/hmap_1_1k.jpg": (1024, 1024, True, False, False),
    "assets/ball/hmap_2_1k.jpg": (1024, 1024, True, False, False),
    "assets/ball/hmap_3_1k.jpg": (1024, 1024, True, False, False),
    "assets/ball/hmap_4_1k.jpg": (1024, 1024, True, False, False),
    "assets/ball/hmap_5_1k.jpg": (1024, 1024, True, False, False),
    "assets/ball/colormap_1k.jpg": (1024, 1024, False, False, False),
    "assets/ball/colormap_2_1k.jpg": (1024, 1024, False, False, False),
    "assets/ball/beammap_1k_1.jpg": (1024, 1024, False, False, False),
    "assets/ball/beammap_1k_2.jpg": (1024, 1024, False, False, False),
    "assets/ball/beammap_1k_3-2.jpg": (1024, 1024, False, False, False),
    "assets/ball/envmap3_1k.jpg": (1024, 1024, False, False, False),
    "assets/ball/nytrik-background_1280x720.png": (720, 1280, False, True, False),
    "assets/ball/nytrik-background-2-1280x720.png": (720, 1280, False, True, False),
    "assets/ball/halo.png": (720, 1280, False, True, False),
    "assets/twister/hmap_2_1k.jpg": (1024, 1024, True, False, False),
    "assets/twister/colormap_1k.jpg": (1024, 1024, False, False, False),
    "assets/twister/nytrik-background_1280x720.png": (720, 1280, False, True, False),
    # loaded by Shared_Create (code/shared-resources.cpp:27-34); only the compositor reads them
    "assets/demo/TPB-logo.png": (720, 1280, False, True, False),
    "assets/demo/tpb_xbox_tp-263x243.png": (243, 263, False, False, False),
}


def default_npz_path():
    return os.environ.get("CKD_ASSETS", os.path.join(_REPO, "oracle", "_ref", "assets.npz"))


def _nearest_resize(arr, new_h, new_w):
    h, w = arr.shape[:2]
    if (h, w) == (new_h, new_w):
        return arr
    yi = (np.arange(new_h, dtype=np.int64) * h) // new_h
    xi = (np.arange(new_w, dtype=np.int64) * w) // new_w
    return np.ascontiguousarray(arr[yi][:, xi])


def _hash_u32(x):
    """integer avalanche hash on uint32 arrays (pure integer arithmetic: identical on every platform / numpy version)"""
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _tri(v, period):
    """integer triangle wave in [0, 255]"""
    v = np.mod(v, period).astype(np.int64)
    half = period // 2
    return (np.where(v < half, v, period - v) * 255 // max(half, 1)).astype(np.int64)


def _synthetic(path, h, w, gray):
    """deterministic procedural stand-in: smooth integer ridges + hash noise (keeps voxel spans / occlusion non-trivial)"""
    seed = sum((i + 1) * b for i, b in enumerate(path.encode())) & 0xFFFFFFFF
    y = np.arange(h, dtype=np.int64)[:, None]
    x = np.arange(w, dtype=np.int64)[None, :]
    p1 = 64 + seed % 97
    p2 = 96 + (seed >> 7) % 131
    ridges = (_tri(x + (seed & 63), p1) * _tri(y + (seed >> 3 & 63), p2)) // 255
    diag = _tri(x + 2 * y, 3 * p1 // 2)
    noise = (_hash_u32((x + y * w + seed).astype(np.uint32)) >> np.uint32(27)).astype(np.int64)  # 0..31
    lum = np.clip((ridges * 5 + diag * 3) // 8 + noise - 16, 0, 255).astype(np.uint8)
    if "tscape-D7" in path:
        lum = (lum.astype(np.uint16) * 150 // 255).astype(np.uint8)  # keep below the view height (code/tunnelscape.cpp:75)
    if "ball/hmap" in path:
        # the reference's ball caster runs off its 1280-pixel rows when height*radius>>8 >= 1280:
        # with ball:Radius clamped to 1920 the stand-in maps stay below 170
        lum = (lum.astype(np.uint16) * 160 // 255).astype(np.uint8)
    if gray:
        return lum
    b = lum
    g = np.roll(lum, w // 7, axis=1)
    r = np.roll(lum, h // 5, axis=0) if h > 1 else np.roll(lum, w // 11, axis=1)
    a = np.full_like(lum, 255)
    if "halo" in path or "background" in path or "blur-map" in path or "foggradient" in path:
        a = np.roll(lum, w // 3, axis=1)
    if "foggradient" in path:
        ramp = (np.arange(w, dtype=np.int64) * 200 // max(w - 1, 1)).astype(np.uint8)[None, :]
        b, g, r = ramp, (ramp // 2 + 20).astype(np.uint8), (255 - ramp).astype(np.uint8) // 3
    rgba = np.stack(np.broadcast_arrays(b, g, r, a), axis=-1)
    return np.ascontiguousarray(rgba).view(np.uint32).reshape(h, w)


class Assets:
    """dict-like: reference image path -> ndarray (uint8 HxW or uint32 HxW as 0xAARRGGBB) for a given output resolution"""

    def __init__(self, res_x=1280, res_y=720, npz_path=None, force_synthetic=False):
        self.res_x, self.res_y = int(res_x), int(res_y)
        self.fx_x, self.fx_y = self.res_x // 2 + 4, self.res_y // 2 + 4
        path = npz_path or default_npz_path()
        self._npz = None
        if not force_synthetic and os.path.isfile(path):
            self._npz = np.load(path)
        self.synthetic = self._npz is None
        self._cache = {}

    def _native(self, path):
        h, w, gray, _, _ = SPEC[path]
        if self._npz is not None and path in self._npz.files:
            return self._npz[path]
        if path == "assets/scape/tscape-C7W-edit.png":
            return _nearest_resize(self._native("assets/scape/C17W-edit.png"), 2048, 2048)
        return _synthetic(path, h, w, gray)

    def __getitem__(self, path):
        if path not in self._cache:
            h, w, gray, out_sized, fx_sized = SPEC[path]
            arr = self._native(path)
            if out_sized:
                arr = _nearest_resize(arr, self.res_y, self.res_x)
            elif fx_sized:
                arr = _nearest_resize(arr, self.fx_y, self.fx_x)
            self._cache[path] = np.ascontiguousarray(arr)
        return self._cache[path]

    def paths(self):
        return list(SPEC.keys())
