"""Input art / maps for the effects, as raw arrays shared byte-for-byte by the CUDA path and the CPU oracle.

The reference decodes its art with DevIL (code/image.cpp:31-73: 32-bit BGRA or 8-bit luminance, upper-left
origin).  Decoding is outside the hot path, so the harness works on pre-decoded arrays:

* ``refdata/assets.npz`` (written by ``oracle/build_ref.py`` from the reference's ``target/assets``) when present,
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

# the compositor's layers and sprites (code/demo.cpp:198-374), in load order.  Output-sized layers are nearest-upscaled with
# the resolution like the effects' own art; sprites and the 1280x568 credit logos stay as they are (the reference places them
# with compile-time constants); the ribbon strip scales with the resolution so that the part-12 strided read (hard-coded
# stride 2160, code/demo.cpp:883) stays inside the image at 4K.
_CREDITS = (
    ["assets/credits/Credits_Tag_Superplek_outlined.png", "assets/credits/Credits_Tag_Comatron_Featuring_Celin_outlined.png",
     "assets/credits/Credits_Tag_Jade_outlined.png", "assets/credits/Credits_Tag_ErnstHot_outlined_new.png"]
    + [f"assets/credits/comatron_anim/comatron_{i}.png" for i in range(1, 6)]
    + [f"assets/credits/animplek/animplek{i}.png" for i in range(5)]
    + [f"assets/credits/jade&nytrik/jade&nytrik{i}.png" for i in range(5)]
    + [f"assets/credits/animhot0/animhot{i}.png" for i in range(5)])
_LAYERS = (
    ["assets/demo/tpb-06-dirty-vignette-1280x720.png"]
    + [f"assets/spikeball/Layer 2023_{i}.png" for i in range(1, 5)]
    + ["assets/spikeball/Vignette_CoolFilmLook.png", "assets/spikeball/Vignette_Layer02_inverted.png",
       "assets/spikeball/SpikeyBall_byPass_BG_Overlay.png", "assets/spikeball/nytrik-TheYearWas_Overlay_LensDirt.jpg"]
    + [f"assets/tunnels/layer 1995_{i}.png" for i in range(1, 5)]
    + [f"assets/tunnels/layer 2006_{i}.png" for i in range(1, 5)]
    + ["assets/tunnels/nytrik-TheYearWas_Overlay_LensDirt.png", "assets/tunnels/Vignette_CoolFilmLook.png",
       "assets/tunnels/Vignette_Layer02_inverted.png", "assets/demo/nytrik-god-layer-720p.png", "assets/scape/revision-logo_white.png",
       "assets/ball/Vignette_Sparta300.png", "assets/greetings/Bokeh_Lens_Dirt_51.png"]
    + [f"assets/greetings/Greetings_Part{i}_BG_Overlay.png" for i in range(1, 5)]
    + ["assets/greetings/Vignette_CoolFilmLook.png", "assets/nautilus/Vignette.png", "assets/nautilus/GlassDirt_Distorted2.png",
       "assets/nautilus/JacquesCousteau_Silhouette2.png", "assets/nautilus/JacquesCousteau1_Silhouette.png",
       "assets/nautilus/JacquesCousteau1_Silhouette_RimMask.png", "assets/nautilus/JacquesCousteau_Silhouette2_RimMask.png",
       "assets/nautilus/JacquesCousteau_Text.png", "assets/closeup/raker-LensDirt5_invert.png", "assets/closeup/VignetteForRaker.png",
       "assets/closeup/Vignette_CoolFilmLook.png", "assets/underwater/LensDirt3_invert.png",
       "assets/underwater/love prism_alpha 1280_720.png"])
_SPRITES = dict(
    [(f"assets/demo/tpb-06-disco-guy/{n}.png", (128, 128)) for n in ("1", "1b", "2", "2b", "3", "3b", "4", "4b")]
    + [("assets/demo/are-we-done-1100x57.png", (57, 1100)), ("assets/closeup/raker_textSmall.png", (115, 624)),
       ("assets/shooting/Lenz.png", (64, 64)), ("assets/demo/GPU-joke.png", (160, 960))])
RIBBONS = "assets/demo/ribbons.png"  # 2160x720, scaled by res_y/720

DEMO_SPEC = {}
for _p in _CREDITS:
    DEMO_SPEC[_p] = (568, 1280, False, False, False)
for _p in _LAYERS:
    DEMO_SPEC[_p] = (720, 1280, False, True, False)
for _p, (_h, _w) in _SPRITES.items():
    DEMO_SPEC[_p] = (_h, _w, False, False, False)
DEMO_SPEC[RIBBONS] = (720, 2160, False, False, False)
SPEC.update(DEMO_SPEC)


def default_npz_path():
    return os.environ.get("CKD_ASSETS", os.path.join(_REPO, "refdata", "assets.npz"))


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
    if "halo" in path or "background" in path or "blur-map" in path or "foggradient" in path or path in DEMO_SPEC:
        a = np.roll(lum, w // 3, axis=1)  # the compositor's layers are mixed by their own alpha: keep it varied
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
            elif path == RIBBONS:
                arr = _nearest_resize(arr, h * self.res_y // 720, w * self.res_y // 720)
            self._cache[path] = np.ascontiguousarray(arr)
        return self._cache[path]

    def paths(self, demo=False):
        """the images of the five effects and Shared_Create; with demo=True also the compositor's (code/demo.cpp:198-374)"""
        return [p for p in SPEC if demo or p not in DEMO_SPEC]

    def drop(self, path):
        """forget the cached array (4K layers are 33 MB each: callers that have handed one over can release it)"""
        self._cache.pop(path, None)
