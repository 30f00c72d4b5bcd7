"""oracle/ref.py -- TEST INFRASTRUCTURE: ctypes driver for oracle/_ref/libckd_ref_<resY>.so.

The library is the *reference itself* (compiled by oracle/build_ref.py); this module only registers the shared
input images, pins the Rocket time and calls the reference's public entry points.  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import it.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
DATA_DIR = os.path.join(os.path.dirname(HERE), "refdata")   # sync/*.track, directors-cut.rocket, assets.npz (oracle/build_ref.py)

ROW_RATE = (170.0 / (60.0 * (170.0 / 174.0))) * 16.0  # code/audio.cpp:18 (= 46.4 rows/s)

EFFECTS = {
    "plasma": 0, "nautilus": 1, "spikey_close": 2, "spikey_distant": 3, "tunnel": 4, "sinuses": 5, "laura": 6,
    "landscape": 7, "tunnelscape": 8, "ball": 9, "twister": 10,
}

# one pinned Rocket row per effect (config 1-3)
CONFIG_ROWS = {
    "plasma": 2600, "nautilus": 5700, "spikey_close": 6800, "spikey_distant": 3600, "tunnel": 4500,
    "sinuses": 7800, "laura": 8900, "landscape": 500, "tunnelscape": 4300, "ball": 1500, "ball_beams": 2060,
    "twister": 2008,
}

BLEND_OPS = {
    "Mix32": 0, "MixOver32": 1, "Add32": 2, "Sub32": 3, "Excl32": 4, "SoftLight32": 5, "SoftLight32A": 6,
    "SoftLight32AA": 7, "Overlay32": 8, "Overlay32A": 9, "Darken32_50": 10, "MulSrc32": 11, "MulSrc32A": 12,
    "MixSrc32": 13, "Fade32": 14,
}
BLIT_OPS = {"BlitSrc32": 0, "BlitSrc32A": 1, "BlitAdd32": 2, "BlitAdd32A": 3}

_U32P = C.POINTER(C.c_uint32)


def lib_path(res_y):
    return os.path.join(REF_DIR, f"libckd_ref_{res_y}.so")


def available(res_y=720):
    return os.path.isfile(lib_path(res_y)) and os.path.isdir(os.path.join(DATA_DIR, "sync"))


def _p32(arr):
    assert arr.dtype == np.uint32 and arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(_U32P)


def aligned_u32(n, pad=64):
    """16-byte aligned uint32 buffer with 'pad' zeroed slack elements after it"""
    raw = np.zeros(n + pad + 4, dtype=np.uint32)
    off = (-raw.ctypes.data % 16) // 4
    return raw[off:off + n]


class Reference:
    _instances = {}

    @classmethod
    def get(cls, res_y=720, assets=None, demo=False):
        """one instance per resolution and process (the reference keeps its state in globals).  demo=True also runs
        Demo_Create (code/demo.cpp:138-374), which needs the compositor's images; it cannot be added afterwards."""
        if res_y not in cls._instances:
            # CKD_REF_DEMO="720[,2160]": create these resolutions with the compositor even when the first caller did not ask
            demo = demo or str(res_y) in os.environ.get("CKD_REF_DEMO", "").split(",")
            cls._instances[res_y] = cls(res_y, assets, demo)
        inst = cls._instances[res_y]
        if demo and not inst.demo:
            raise RuntimeError("the reference was already created without the compositor in this process")
        return inst

    def __init__(self, res_y=720, assets=None, demo=False):
        self.lib = C.CDLL(lib_path(res_y))
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        L.ref_register_image.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
        L.ref_create.argtypes = [C.c_char_p]
        L.ref_set_time.argtypes = [C.c_double]
        L.ref_row.restype = C.c_double
        L.ref_track.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_track.restype = C.c_double
        L.ref_draw.argtypes = [C.c_int, _U32P, C.c_float, C.c_float]
        L.ref_fxmap.argtypes = [C.c_int]
        L.ref_fxmap.restype = _U32P
        L.ref_render_target.argtypes = [C.c_int]
        L.ref_render_target.restype = _U32P
        L.ref_cos_lut.restype = C.POINTER(C.c_float)
        L.ref_fast_cos_tab.restype = C.POINTER(C.c_double)
        L.ref_fx_blit_2x2.argtypes = [_U32P, _U32P]
        L.ref_polar_blit.argtypes = [_U32P, _U32P, C.c_int]
        L.ref_polar_blit_a.argtypes = [_U32P, _U32P, C.c_int]
        L.ref_polar_blit_2x2.argtypes = [_U32P, _U32P, C.c_int]
        L.ref_fx_test_pattern.argtypes = [_U32P]
        L.ref_ball_background.restype = C.c_void_p
        for name in ("ref_old_blur_h", "ref_old_blur_v", "ref_old_blur"):
            getattr(L, name).argtypes = [_U32P, _U32P, C.c_uint, C.c_uint, C.c_float]
        for name in ("ref_new_blur_h", "ref_new_blur_v", "ref_new_blur"):
            getattr(L, name).argtypes = [_U32P, _U32P, C.c_uint, C.c_uint, C.c_float, C.c_float, C.c_uint]
        L.ref_box_blur_scale.argtypes = [C.c_float]
        L.ref_box_blur_scale.restype = C.c_float
        L.ref_memset32.argtypes = [_U32P, C.c_int, C.c_size_t]
        L.ref_tape_warp.argtypes = [_U32P, _U32P, C.c_uint, C.c_uint, C.c_float, C.c_float]
        L.ref_blend.argtypes = [C.c_int, _U32P, _U32P, C.c_uint, C.c_float, C.c_uint]
        L.ref_blit.argtypes = [C.c_int, _U32P, _U32P, C.c_uint, C.c_uint, C.c_uint, C.c_float]
        L.ref_mix_src_s.argtypes = [_U32P, _U32P, C.c_uint, C.c_uint, C.c_uint]
        for name in ("ref_lutcosf", "ref_lutsinf", "ref_q3_rsqrtf2", "ref_rsqrt_ss", "ref_expf"):
            getattr(L, name).argtypes = [C.c_float]
            getattr(L, name).restype = C.c_float
        L.ref_fastcosf.argtypes = [C.c_double]
        L.ref_fastcosf.restype = C.c_float
        for name in ("ref_powf", "ref_atan2f"):
            getattr(L, name).argtypes = [C.c_float, C.c_float]
            getattr(L, name).restype = C.c_float
        L.ref_rsqrt_scan.argtypes = [_U32P, C.c_uint]
        L.ref_log_ps.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_exp_ps.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_gamma_pixels.argtypes = [C.POINTER(C.c_float), C.c_float, _U32P, C.c_uint]
        L.ref_to_pixels_noconv.argtypes = [C.POINTER(C.c_float), _U32P, C.c_uint]
        L.ref_cspan16.argtypes = [_U32P, C.c_int, C.c_uint, C.c_uint, C.c_uint32, C.c_uint32]
        L.ref_bsamp8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint]
        L.ref_bsamp8.restype = C.c_uint
        L.ref_bsamp32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint]
        L.ref_bsamp32.restype = C.c_uint32

        self.res_x, self.res_y = L.ref_res_x(), L.ref_res_y()
        self.fx_x, self.fx_y = L.ref_fxmap_res_x(), L.ref_fxmap_res_y()
        assert self.res_y == res_y

        if assets is None:
            from cookiedough_b200.assets import Assets
            assets = Assets(self.res_x, self.res_y)
        self.assets = assets
        self.demo = bool(demo)
        L.ref_create_demo.argtypes = [C.c_char_p]
        L.ref_demo_draw.argtypes = [_U32P, C.c_float, C.c_float]
        self._keep = []
        for path in assets.paths(demo=self.demo):
            arr = assets[path]
            self._keep.append(arr)
            L.ref_register_image(path.encode(), arr.ctypes.data, arr.nbytes)
        rc = (L.ref_create_demo if self.demo else L.ref_create)(DATA_DIR.encode())
        if rc != 0:
            raise RuntimeError(f"ref_create failed ({rc}): {L.ref_last_error().decode()}")
        self._keep = []  # Image_Load* copied the pixels

    # -- timeline -------------------------------------------------------------------------------
    def set_row(self, row):
        self.set_time(row / ROW_RATE)

    def set_time(self, seconds):
        self.time = float(seconds)
        self.lib.ref_set_time(self.time)

    def track(self, name):
        return self.lib.ref_track(DATA_DIR.encode(), name.encode())

    # -- effects --------------------------------------------------------------------------------
    def frame(self):
        return aligned_u32(self.res_x * self.res_y, pad=self.res_x * 2).reshape(self.res_y, self.res_x)

    def draw(self, effect, out=None, delta=1.6667):
        """calls X_Draw(pDest, time, delta) of the reference at the pinned time"""
        if out is None:
            out = self.frame()
        rc = self.lib.ref_draw(EFFECTS[effect], _p32(out), C.c_float(self.time), C.c_float(delta))
        assert rc == 0
        return out

    def demo_draw(self, out=None, delta=1.6667):
        """Demo_Draw(pDest, time, delta) (code/demo.cpp:469-1023) at the pinned time; it advances Rocket itself"""
        assert self.demo
        if out is None:
            out = self.frame()
        self.lib.ref_demo_draw(_p32(out), C.c_float(self.time), C.c_float(delta))
        return out

    def fxmap(self, i):
        p = self.lib.ref_fxmap(i)
        return np.ctypeslib.as_array(p, shape=(self.fx_y, self.fx_x))

    def render_target(self, i):
        p = self.lib.ref_render_target(i)
        return np.ctypeslib.as_array(p, shape=(self.res_y, self.res_x))

    def cos_lut(self):
        return np.ctypeslib.as_array(self.lib.ref_cos_lut(), shape=(2049,)).copy()

    # -- post ops (operate on caller buffers; dst may alias src) ----------------------------------
    def fx_blit_2x2(self, dst, src):
        self.lib.ref_fx_blit_2x2(_p32(dst), _p32(src))

    def polar_blit(self, dst, src, inverse=False, alpha=False):
        (self.lib.ref_polar_blit_a if alpha else self.lib.ref_polar_blit)(_p32(dst), _p32(src), int(inverse))

    def polar_blit_2x2(self, dst, src, inverse=False):
        """Polar_Blit_2x2 on FX-map sized buffers (oracle patched to stay inside them, build_ref.py P2/P2b)"""
        self.lib.ref_polar_blit_2x2(_p32(dst), _p32(src), int(inverse))

    def fx_test_pattern(self, dst):
        self.lib.ref_fx_test_pattern(_p32(dst))

    def ball_background(self):
        addr = self.lib.ref_ball_background()
        n = self.res_x * self.res_y
        return np.frombuffer((C.c_uint32 * n).from_address(addr), dtype=np.uint32).copy()

    def old_blur(self, kind, dst, src, w, h, strength):
        fn = {"h": self.lib.ref_old_blur_h, "v": self.lib.ref_old_blur_v, "hv": self.lib.ref_old_blur}[kind]
        fn(_p32(dst), _p32(src), w, h, C.c_float(strength))

    def new_blur(self, kind, dst, src, w, h, strength, gain, passes):
        fn = {"h": self.lib.ref_new_blur_h, "v": self.lib.ref_new_blur_v, "hv": self.lib.ref_new_blur}[kind]
        fn(_p32(dst), _p32(src), w, h, C.c_float(strength), C.c_float(gain), passes)

    def blend(self, op, dst, src, fparam=0.0, uparam=0, n=None):
        rc = self.lib.ref_blend(BLEND_OPS[op], _p32(dst), _p32(src), dst.size if n is None else n, C.c_float(fparam), C.c_uint(uparam))
        assert rc == 0

    def blit(self, op, dst, src, dest_res_x, src_res_x, y_res, alpha=1.0):
        rc = self.lib.ref_blit(BLIT_OPS[op], _p32(dst), _p32(src), dest_res_x, src_res_x, y_res, C.c_float(alpha))
        assert rc == 0

    def mix_src_s(self, dst, src, dest_res_x, dest_res_y, src_stride):
        self.lib.ref_mix_src_s(_p32(dst), _p32(src), dest_res_x, dest_res_y, src_stride)

    def memset32(self, dst, value, n):
        self.lib.ref_memset32(_p32(dst), C.c_int(int(np.uint32(value).astype(np.int32))), n)

    def tape_warp(self, dst, src, w, h, strength, speed):
        self.lib.ref_tape_warp(_p32(dst), _p32(src), w, h, C.c_float(strength), C.c_float(speed))

    def rsqrt_table(self, stride=1):
        n = 2 * ((1 << 23) // stride)
        out = np.zeros(n, dtype=np.uint32)
        self.lib.ref_rsqrt_scan(_p32(out), stride)
        return out
