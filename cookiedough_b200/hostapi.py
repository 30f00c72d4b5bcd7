"""ctypes driver of the C++ host layer (include/ckd_host.h): the reference's own X_Create / X_Draw(uint32_t *pDest, float
time, float delta) entry points with HOST buffers.  bench.py's end-to-end leg and the drop-in tests go through this."""
import ctypes as C
import os

import numpy as np

from . import capi

ROW_RATE = (170.0 / (60.0 * (170.0 / 174.0))) * 16.0  # code/audio.cpp:18

EFFECT_IDS = {"plasma": 0, "nautilus": 1, "spikey_close": 2, "spikey_distant": 3, "tunnel": 4, "sinuses": 5, "laura": 6,
              "landscape": 7, "tunnelscape": 8, "ball": 9, "twister": 10}
POST_IDS = {"Fx_Blit_2x2": 0, "Polar_Blit": 1, "Polar_BlitA": 2, "HorizontalBoxBlur32": 3, "VerticalBoxBlur32": 4, "BoxBlur32": 5,
            "BoxBlur_32": 6, "MixSrc32": 7, "SoftLight32": 8, "TapeWarp32": 9, "Polar_Blit_2x2": 10, "FxBlitter_DrawTestPattern": 11,
            "BlitSrc32": 12, "BlitSrc32A": 13, "BlitAdd32": 14, "BlitAdd32A": 15, "MixSrc32S": 16, "memset32": 17}
MODULE_IDS = {"Polar": 0, "BoxBlur": 1, "FxBlitter": 2, "Shared": 3}
GLOBAL_IDS = {"g_pFxMap": 0, "g_renderTarget": 4, "g_pNytrikTPB": 8, "g_pXboxLogoTPB": 9, "g_gradientUnp16": 10, "Ball_GetBackground": 11}

_U32P = C.POINTER(C.c_uint32)


def _lib():
    L = capi.load()
    if getattr(L, "_host_bound", False):
        return L
    L.ckdhost_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.ckdhost_launch.argtypes = []
    L.ckdhost_register_image.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ckdhost_last_error.restype = C.c_char_p
    L.ckdhost_context.restype = C.c_void_p
    L.ckdhost_set_time.argtypes = [C.c_double]
    L.ckdhost_rocket_open.argtypes = [C.c_char_p]
    L.ckdhost_track.argtypes = [C.c_char_p]
    L.ckdhost_track.restype = C.c_double
    L.ckdhost_track_i.argtypes = [C.c_char_p]
    L.ckdhost_draw.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_float]
    L.ckdhost_post.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_float, C.c_float, C.c_uint]
    L.ckdhost_release_image.argtypes = [C.c_char_p]
    L.ckdhost_effect_create.argtypes = [C.c_int]
    L.ckdhost_set_asset_root.argtypes = [C.c_char_p]
    L.ckdhost_module.argtypes = [C.c_int, C.c_int]
    L.ckdhost_set_readback_bands.argtypes = [C.c_int]
    L.ckdhost_global.argtypes = [C.c_int]
    L.ckdhost_global.restype = C.c_void_p
    L.ckdhost_demo_create.argtypes = []
    L.ckdhost_demo_draw.argtypes = [C.c_void_p, C.c_double, C.c_float]
    L.ckdhost_demo_destroy.argtypes = []
    L.ckdhost_timeline_render.argtypes = [C.POINTER(C.c_double), C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_int,
                                          C.POINTER(C.c_void_p), C.c_uint, C.c_ulonglong, C.c_float, C.c_uint, C.c_uint]
    L.ckdhost_launch_count.restype = C.c_ulonglong
    L.ckdhost_timeline_owner.argtypes = [C.c_uint, C.c_uint, C.c_uint]
    L.ckdhost_timeline_owner.restype = C.c_uint
    L.ckdhost_timeline_default_skip.argtypes = [C.c_uint]
    L.ckdhost_timeline_default_skip.restype = C.c_uint
    L.ckdhost_fastcos.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.ckdhost_fast_cos_tab.restype = C.POINTER(C.c_double)
    L._host_bound = True
    return L


def default_rocket_source():
    """the demo's Rocket project: the XML shipped with the reference when refdata/ holds a copy (oracle/build_ref.py)"""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.environ.get("CKD_ROCKET", os.path.join(here, "refdata", "directors-cut.rocket"))


class RocketOnly:
    """the host layer's GNU Rocket reader without a GPU context (CPU-side tests)"""

    def __init__(self, source):
        self.L = _lib()
        if self.L.ckdhost_rocket_open(str(source).encode()) != 0:
            raise capi.CkdError(self.L.ckdhost_last_error().decode())

    def set_time(self, seconds):
        return self.L.ckdhost_set_time(float(seconds))

    def set_row(self, row):
        return self.set_time(row / ROW_RATE)

    def track(self, name):
        return self.L.ckdhost_track(name.encode())

    def track_i(self, name):
        return self.L.ckdhost_track_i(name.encode())


class Host:
    """CkdHost_Create + Rocket::Launch + the five X_Create (code/main.cpp:263-279, code/demo.cpp:140-148)"""

    def __init__(self, res_x, res_y, device, assets, rocket_source=None, demo=False):
        """demo=True: Demo_Create (code/demo.cpp:138-367) instead of the bare effects: also loads the compositor's art"""
        self.L = _lib()
        self.res_x, self.res_y = res_x, res_y
        self.demo = bool(demo)
        source = rocket_source or default_rocket_source()
        if self.L.ckdhost_create(res_x, res_y, device, str(source).encode()) != 0:
            raise capi.CkdError(self.L.ckdhost_last_error().decode())
        paths = list(capi.IMAGE_SLOTS)
        if self.demo:
            paths += [p for p in assets.paths(demo=True) if p not in capi.IMAGE_SLOTS]
        for path in paths:
            arr = assets[path]
            self.L.ckdhost_register_image(path.encode(), arr.ctypes.data, arr.shape[1], arr.shape[0], 1 if arr.dtype == np.uint8 else 4)
            if self.demo and arr.nbytes > (8 << 20):
                assets.drop(path)  # the host layer copied it; 4K layers are 33 MB each
        rc = self.L.ckdhost_demo_create() if self.demo else self.L.ckdhost_launch()
        if rc != 0:
            raise capi.CkdError(f"host launch failed ({rc}): {self.L.ckdhost_last_error().decode()}")
        self.ctx_handle = self.L.ckdhost_context()
        self.time = 0.0

    def register_image(self, path, arr):
        """CkdHost_RegisterImage: pre-decoded pixels (uint32 BGRA or uint8 L8, HxW) under the path the reference loads them by"""
        arr = np.ascontiguousarray(arr)
        self.L.ckdhost_register_image(path.encode(), arr.ctypes.data, arr.shape[1], arr.shape[0], 1 if arr.dtype == np.uint8 else 4)

    def reload_from_files(self, effect_module, paths, asset_root):
        """drops the registered copies of 'paths' and runs the module's X_Create again, which then decodes the files under
        asset_root with the host layer's own PNG/JPEG decoders -- the reference's Image_Load32/Image_Load8 route"""
        for path in paths:
            self.L.ckdhost_release_image(path.encode())
        self.L.ckdhost_set_asset_root(str(asset_root).encode())
        rc = self.L.ckdhost_effect_create({"twister": 0, "landscape": 1, "ball": 2, "tunnelscape": 3, "shadertoy": 4}[effect_module])
        if rc != 0:
            raise capi.CkdError(f"{effect_module}: {self.L.ckdhost_last_error().decode()}")

    def context(self):
        """a capi.Context view of the host layer's ckd_ctx (not owning)"""
        ctx = capi.Context.__new__(capi.Context)
        ctx.L = self.L
        ctx.h = C.c_void_p(self.ctx_handle)
        ctx.res_x, ctx.res_y = self.res_x, self.res_y
        ctx.fx_x, ctx.fx_y = self.res_x // 2 + 4, self.res_y // 2 + 4
        ctx._user = []
        ctx.close = lambda: None
        return ctx

    def set_time(self, seconds):
        self.time = float(seconds)
        return self.L.ckdhost_set_time(self.time)

    def set_row(self, row):
        return self.set_time(row / ROW_RATE)

    def track(self, name):
        return self.L.ckdhost_track(name.encode())

    def draw(self, effect, out, delta=1.6667):
        """X_Draw(pDest, time, delta) into the caller's host buffer (numpy uint32, or a raw address)"""
        ptr = out if isinstance(out, int) else out.ctypes.data
        rc = self.L.ckdhost_draw(EFFECT_IDS[effect], C.c_void_p(ptr), C.c_float(self.time), C.c_float(delta))
        if rc != 0:
            raise capi.CkdError(f"{effect}: {self.L.ckdhost_last_error().decode()}")
        return out

    def demo_draw(self, out, seconds=None, delta=1.6667):
        """Demo_Draw(pDest, time, delta) at 'seconds' (default: the time last set); -> False when the demo is over"""
        assert self.demo
        if seconds is not None:
            self.time = float(seconds)
        ptr = out if isinstance(out, int) else out.ctypes.data
        rc = self.L.ckdhost_demo_draw(C.c_void_p(ptr), C.c_double(self.time), C.c_float(delta))
        if rc < 0:
            raise capi.CkdError(f"Demo_Draw: {self.L.ckdhost_last_error().decode()}")
        return rc == 1

    def timeline_render(self, times, rank=0, world=1, gather=None, passes=1, pop_mode=0, host_ring=None, seq_base=0, delta=1.6667, collector_skip=None, lanes=1):
        """CkdTimeline_Render: Demo_Draw for the frames i % world == rank of `times`, each published to `gather` (capi.Gather);
        rank 0 also consumes every frame in order (pop_mode: capi.GATHER_CHECKSUM / GATHER_TO_HOST into host_ring, a list of
        page-locked buffer addresses, or into the open sink when host_ring is None).  lanes=2: this rank's frames alternate between
        two contexts with their own streams (two frames in flight side by side)"""
        assert self.demo
        arr = (C.c_double * len(times))(*times)
        ring = (C.c_void_p * len(host_ring))(*host_ring) if host_ring else None
        if collector_skip is None:      # the library's default for this many GPUs (only matters with a gather: rank 0 collects)
            collector_skip = self.L.ckdhost_timeline_default_skip(world) if gather is not None else 1
        rc = self.L.ckdhost_timeline_render(arr, len(times), passes, rank, world, gather.g if gather is not None else None, pop_mode,
                                            ring, len(host_ring) if host_ring else 0, seq_base, C.c_float(delta), collector_skip, lanes)
        if rc != 0:
            raise capi.CkdError(f"CkdTimeline_Render: {self.L.ckdhost_last_error().decode()}")

    def launch_count(self):
        """kernels launched so far by every lane of the host layer (CkdHost_LaunchCount)"""
        return int(self.L.ckdhost_launch_count())

    def fastcos(self, x, sine=False):
        """InitializeFastCosine + fastcosf / fastsinf over an array (ckd_host.h)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(x.shape, dtype=np.float32)
        if self.L.ckdhost_fastcos(out.ctypes.data, x.ctypes.data, x.size, int(bool(sine))) != 0:
            raise capi.CkdError(f"fastcosf: {self.L.ckdhost_last_error().decode()}")
        return out

    def fast_cos_tab(self):
        return np.ctypeslib.as_array(self.L.ckdhost_fast_cos_tab(), shape=(1025,)).copy()

    def post(self, op, dst, src, a=0, b=0, f0=0.0, f1=0.0, u=0):
        rc = self.L.ckdhost_post(POST_IDS[op], dst.ctypes.data, src.ctypes.data if src is not None else None, a, b, C.c_float(f0), C.c_float(f1), u)
        if rc != 0:
            raise capi.CkdError(f"{op}: {self.L.ckdhost_last_error().decode()}")

    def module(self, name, create=True):
        """Polar/BoxBlur/FxBlitter/Shared _Create (-> bool) or _Destroy"""
        return self.L.ckdhost_module(MODULE_IDS[name], int(bool(create))) == 0

    def global_array(self, name, index=0, shape=None, dtype=np.uint32):
        """numpy view of one of the reference's globals (g_pFxMap[i], g_renderTarget[i], ...); None while it is null"""
        addr = self.L.ckdhost_global(GLOBAL_IDS[name] + index)
        if not addr:
            return None
        n = int(np.prod(shape))
        buf = (C.c_uint8*(n*np.dtype(dtype).itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def pin(self, out):
        """CkdHost_PinFrameBuffer: page-lock a caller-owned numpy frame in place"""
        if self.L.ckdhost_pin_frame_buffer(C.c_void_p(out.ctypes.data)) != 0:
            raise capi.CkdError(self.L.ckdhost_last_error().decode())

    def unpin(self, out):
        self.L.ckdhost_unpin_frame_buffer(C.c_void_p(out.ctypes.data))

    def set_readback_bands(self, bands):
        """CkdHost_SetReadbackBands: -1 automatic, 0 off, n >= 2 row bands for the streamed read-back of a synchronous X_Draw"""
        self.L.ckdhost_set_readback_bands(int(bands))

    def set_pipelined(self, enabled):
        """frame pipelining (CkdHost_SetPipelined): X_Draw returns once enqueued; call flush() before reading the buffers"""
        self.L.ckdhost_set_pipelined(int(bool(enabled)))

    def flush(self):
        self.L.ckdhost_flush()

    def close(self):
        if self.demo:
            self.L.ckdhost_demo_destroy()
        else:
            self.L.ckdhost_destroy()
