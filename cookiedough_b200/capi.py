"""ctypes binding of the C ABI in include/ckd.h (libckd_b200.so).  Plumbing for tests and bench.py.

There is no fallback: if the shared library is missing or no B200 is present, loading / creating a context raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libckd_b200.so")

_U32P = C.POINTER(C.c_uint32)


class CkdError(RuntimeError):
    pass


def _struct(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


_F, _I = C.c_float, C.c_int

PlasmaParams = _struct("PlasmaParams", [(n, _F) for n in ("speed", "hue", "gamma", "desaturation")])
NautilusParams = _struct("NautilusParams", [(n, _F) for n in ("roll", "hue", "speed", "desaturation", "blur")])
SpikeyParams = _struct("SpikeyParams", [
    ("speed", _F), ("roll", _F), ("specular", _F), ("desaturation", _F), ("hue", _F), ("gamma", _F), ("warmup", _F),
    ("dist_x", _F), ("dist_y", _F), ("dist_z", _F), ("close_x", _F), ("close_y", _F), ("close_z", _F),
    ("close_z_scale", _F), ("close_normal_grain", _F), ("close_scale", _F), ("close_rim", _I), ("close_aspect_mul", _I),
    ("mix_blur_map", _F), ("mix_blur", _F), ("mix_map_blur", _F), ("mix_blur_opacity", _F)])
TunnelParams = _struct("TunnelParams", [
    ("boxy", _F), ("flower_scale", _F), ("flower_freq", _F), ("flower_phase", _F), ("speed", _F), ("roll", _F),
    ("pitch", _F), ("radius", _F), ("mul_u", _F), ("mul_v", _F), ("lit_tiles", _I), ("lit_blur", _F), ("fog1", _F), ("fog2", _F)])
SinusesParams = _struct("SinusesParams", [(n, _F) for n in ("specular", "roll", "speed", "offs_x", "gamma", "hue", "desaturation")])
LauraParams = _struct("LauraParams", [(n, _F) for n in ("speed", "yaw", "pitch", "roll", "hue", "saturate")])
LandscapeParams = _struct("LandscapeParams", [(n, _F) for n in (
    "forward", "tilt", "warp_speed", "warp_strength", "pad_tilt", "view_angle", "strafe_x", "strafe_y", "pad_move_x", "pad_move_y")])
TunnelscapeParams = _struct("TunnelscapeParams", [(n, _F) for n in ("step_u", "step_v", "speed", "blur")])
BallParams = _struct("BallParams", [
    ("blur", _F), ("radius", _F), ("ray_length", _I), ("spikes", _I), ("has_beams", _I), ("base_shape_index", _I),
    ("speed", _F), ("beam_atten", _I), ("beam_alpha_min", _F), ("rotate_offs_x", _F), ("rotate_offs_y", _F),
    ("beams1", _F), ("beams2", _F), ("beams3", _F), ("low_beams", _I)])
TwisterParams = _struct("TwisterParams", [(n, _F) for n in ("speed", "shear_speed", "blur")])

# Rocket track behind every struct field (code/shadertoy.cpp:104-171, landscape.cpp:216-219, tunnelscape.cpp:156-159,
# ball.cpp:427-441, torus-twister.cpp:155-157); int fields are read with Rocket::geti = int(roundf(float(v))) (rocket.h:27-29)
TRACKS = {
    "plasma": (PlasmaParams, {"speed": "plasma:Speed", "hue": "plasma:Hue", "gamma": "plasma:Gamma", "desaturation": "plasma:Desaturation"}),
    "nautilus": (NautilusParams, {"roll": "nautilus:Roll", "hue": "nautilus:Hue", "speed": "nautilus:Speed", "desaturation": "nautilus:Desat", "blur": "nautilus:Blur"}),
    "spikey": (SpikeyParams, {
        "speed": "spike:Speed", "roll": "spike:Roll", "specular": "spike:Specular", "desaturation": "spike:Desaturation",
        "hue": "spike:Hue", "gamma": "spike:Gamma", "warmup": "distSpike:Warmup",
        "dist_x": "distSpike:xOffs", "dist_y": "distSpike:yOffs", "dist_z": "distSpike:zOffs",
        "close_x": "closeSpike:xOffs", "close_y": "closeSpike:yOffs", "close_z": "closeSpike:zOffs",
        "close_z_scale": "closeSpike:zOffsScale", "close_normal_grain": "closeSpike:NormalGrain", "close_scale": "closeSpike:Scale",
        "close_rim": "closeSpike:Rim", "close_aspect_mul": "closeSpike:AspectMul", "mix_blur_map": "closeSpike:MixBlurMap",
        "mix_blur": "closeSpike:MixBlur", "mix_map_blur": "closeSpike:MixMapBlur", "mix_blur_opacity": "closeSpike:MixBlurOpacity"}),
    "tunnel": (TunnelParams, {
        "boxy": "tunnel:Boxy", "flower_scale": "tunnel:FlowerScale", "flower_freq": "tunnel:FlowerFreq", "flower_phase": "tunnel:FlowerPhase",
        "speed": "tunnel:Speed", "roll": "tunnel:Roll", "pitch": "tunnel:Pitch", "radius": "tunnel:Radius", "mul_u": "tunnel:MulU",
        "mul_v": "tunnel:MulV", "lit_tiles": "tunnel:LitTiles", "lit_blur": "tunnel:LitBlur", "fog1": "tunnel:Fog1", "fog2": "tunnel:Fog2"}),
    "sinuses": (SinusesParams, {
        "specular": "sinusesTunnel:Specular", "roll": "sinusesTunnel:Roll", "speed": "sinusesTunnel:Speed", "offs_x": "sinusesTunnel:OffsX",
        "gamma": "sinusesTunnel:Gamma", "hue": "sinusesTunnel:Hue", "desaturation": "sinusesTunnel:Desaturation"}),
    "laura": (LauraParams, {"speed": "laura:Speed", "yaw": "laura:Yaw", "pitch": "laura:Pitch", "roll": "laura:Roll", "hue": "laura:Hue", "saturate": "laura:Saturate"}),
    "landscape": (LandscapeParams, {"forward": "voxelScape:Forward", "tilt": "voxelScape:Tilt", "warp_speed": "voxelScape:WarpSpeed", "warp_strength": "voxelScape:WarpStrength"}),
    "tunnelscape": (TunnelscapeParams, {"step_u": "starsTunnel:stepU", "step_v": "starsTunnel:stepV", "speed": "starsTunnel:Speed", "blur": "starsTunnel:Blur"}),
    "ball": (BallParams, {
        "blur": "ball:Blur", "radius": "ball:Radius", "ray_length": "ball:RayLength", "spikes": "ball:Spikes", "has_beams": "ball:HasBeams",
        "base_shape_index": "ball:BaseShapeIndex", "speed": "ball:Speed", "beam_atten": "ball:BeamAttenuation", "beam_alpha_min": "ball:BeamAlphaMin",
        "rotate_offs_x": "ball:RotateOffsX", "rotate_offs_y": "ball:RotateOffsY", "beams1": "ball:Beams1", "beams2": "ball:Beams2",
        "beams3": "ball:Beams3", "low_beams": "ball:BallLowBeams"}),
    "twister": (TwisterParams, {"speed": "twister:Speed", "shear_speed": "twister::ShearSpeed", "blur": "twister:Blur"}),
}

IMAGE_SLOTS = {
    "assets/shadertoy/nytrik-hextexture.png": 0,
    "assets/shadertoy/nytrik-hextexture-fx.png": 1,
    "assets/shadertoy/close-up-blur-map-1.png": 2,
    "assets/shadertoy/close-up-blur-map-2.png": 3,
    "assets/scape/D17.png": 4,
    "assets/scape/C17W-edit.png": 5,
    "assets/scape/foggradient.jpg": (6, 9),
    "assets/scape/tscape-D7-edit.png": 7,
    "assets/scape/tscape-C7W-edit.png": 8,
    "assets/ball/hmap_1_1k.jpg": 10,
    "assets/ball/hmap_4_1k.jpg": 11,
    "assets/ball/hmap_2_1k.jpg": 12,
    "assets/ball/hmap_3_1k.jpg": 13,
    "assets/ball/hmap_5_1k.jpg": 14,
    "assets/ball/colormap_1k.jpg": 15,
    "assets/ball/colormap_2_1k.jpg": 16,
    "assets/ball/beammap_1k_1.jpg": 17,
    "assets/ball/beammap_1k_2.jpg": 18,
    "assets/ball/beammap_1k_3-2.jpg": 19,
    "assets/ball/envmap3_1k.jpg": 20,
    "assets/ball/nytrik-background_1280x720.png": 21,
    "assets/ball/nytrik-background-2-1280x720.png": 22,
    "assets/ball/halo.png": 23,
    "assets/twister/hmap_2_1k.jpg": 24,
    "assets/twister/colormap_1k.jpg": 25,
    "assets/twister/nytrik-background_1280x720.png": 26,
}

BLEND_OPS = {
    "Mix32": 0, "MixOver32": 1, "Add32": 2, "Sub32": 3, "Excl32": 4, "SoftLight32": 5, "SoftLight32A": 6,
    "SoftLight32AA": 7, "Overlay32": 8, "Overlay32A": 9, "Darken32_50": 10, "MulSrc32": 11, "MulSrc32A": 12,
    "MixSrc32": 13, "Fade32": 14,
}
BLIT_OPS = {"BlitSrc32": 0, "BlitSrc32A": 1, "BlitAdd32": 2, "BlitAdd32A": 3}



class BlendStep(C.Structure):  # ckd_blend_step
    _fields_ = [("op", C.c_int), ("d_src", C.c_void_p), ("f_param", C.c_float), ("u_param", C.c_uint)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint), ("total_ms", C.c_double), ("algo_bytes", C.c_double)]


_lib = None


def load():
    """dlopen libckd_b200.so (raises if it has not been built: there is no other implementation to fall back to)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CkdError(f"{LIB_PATH} is missing: run `python -m cookiedough_b200.build` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    VP, SZ, U, F = C.c_void_p, C.c_size_t, C.c_uint, C.c_float
    sig = {
        "ckd_create": ([C.POINTER(VP), _I, _I, _I], _I),
        "ckd_destroy": ([VP], None),
        "ckd_last_error": ([], C.c_char_p),
        "ckd_version": ([], C.c_char_p),
        "ckd_set_stream": ([VP, VP], _I),
        "ckd_sync": ([VP], _I),
        "ckd_res_x": ([VP], _I), "ckd_res_y": ([VP], _I), "ckd_fxmap_res_x": ([VP], _I), "ckd_fxmap_res_y": ([VP], _I),
        "ckd_frame": ([VP], VP), "ckd_fxmap": ([VP, _I], VP), "ckd_render_target": ([VP, _I], VP),
        "ckd_malloc": ([VP, C.POINTER(VP), SZ], _I), "ckd_free": ([VP, VP], _I),
        "ckd_malloc_host": ([C.POINTER(VP), SZ], _I), "ckd_free_host": ([VP], _I),
        "ckd_upload": ([VP, VP, VP, SZ], _I), "ckd_download": ([VP, VP, VP, SZ], _I), "ckd_copy": ([VP, VP, VP, SZ], _I), "ckd_pin_host": ([VP, SZ], _I), "ckd_unpin_host": ([VP], _I),
        "ckd_timer_start": ([VP], _I), "ckd_timer_stop_ms": ([VP, C.POINTER(F)], _I),
        "ckd_set_cos_lut": ([VP, C.POINTER(F)], _I),
        "ckd_set_rsqrt_table": ([VP, _U32P, _I], _I),
        "ckd_get_rsqrt_table": ([VP, _U32P, SZ, C.POINTER(_I), C.POINTER(SZ)], _I),
        "ckd_set_polar_maps": ([VP, VP, VP], _I), "ckd_get_polar_maps": ([VP, VP, VP], _I),
        "ckd_set_image": ([VP, _I, VP, _I, _I, _I], _I), "ckd_get_image": ([VP, _I], VP),
        "ckd_fx_blit_2x2": ([VP, VP, VP], _I),
        "ckd_polar_blit": ([VP, VP, VP, _I], _I), "ckd_polar_blit_a": ([VP, VP, VP, _I], _I), "ckd_polar_blit_2x2": ([VP, VP, VP, _I], _I),
        "ckd_old_blur_h": ([VP, VP, VP, U, U, F], _I), "ckd_old_blur_v": ([VP, VP, VP, U, U, F], _I), "ckd_old_blur": ([VP, VP, VP, U, U, F], _I),
        "ckd_box_blur_scale": ([F], F),
        "ckd_new_blur_h": ([VP, VP, VP, U, U, F, F, U], _I), "ckd_new_blur_v": ([VP, VP, VP, U, U, F, F, U], _I), "ckd_new_blur": ([VP, VP, VP, U, U, F, F, U], _I),
        "ckd_blend": ([VP, _I, VP, VP, U, F, U], _I), "ckd_blend_chain": ([VP, VP, VP, U, U], _I),
        "ckd_blit": ([VP, _I, VP, VP, U, U, U, F], _I),
        "ckd_mix_src_s": ([VP, VP, VP, U, U, U], _I),
        "ckd_memset32": ([VP, VP, C.c_uint32, SZ], _I),
        "ckd_tape_warp": ([VP, VP, VP, U, U, F, F], _I),
        "ckd_plasma_draw": ([VP, C.POINTER(PlasmaParams), F, VP], _I),
        "ckd_nautilus_draw": ([VP, C.POINTER(NautilusParams), F, VP], _I),
        "ckd_spikey_draw": ([VP, C.POINTER(SpikeyParams), F, _I, VP], _I),
        "ckd_tunnel_draw": ([VP, C.POINTER(TunnelParams), F, VP], _I),
        "ckd_sinuses_draw": ([VP, C.POINTER(SinusesParams), F, VP], _I),
        "ckd_laura_draw": ([VP, C.POINTER(LauraParams), F, VP], _I),
        "ckd_landscape_draw": ([VP, C.POINTER(LandscapeParams), F, VP], _I),
        "ckd_tunnelscape_draw": ([VP, C.POINTER(TunnelscapeParams), F, VP], _I),
        "ckd_ball_draw": ([VP, C.POINTER(BallParams), F, VP], _I),
        "ckd_ball_beam_tail": ([VP, VP, _I, _I, _I, C.c_uint32, F, _I], _I),
        "ckd_twister_draw": ([VP, C.POINTER(TwisterParams), F, VP], _I),
        "ckd_launch_count": ([VP], C.c_ulonglong),
        "ckd_set_fast_cos_table": ([VP, VP], _I), "ckd_get_fast_cos_table": ([VP, VP], _I),
        "ckd_fastcos": ([VP, VP, VP, SZ, _I], _I),
        "ckd_set_frame_independent": ([VP, _I], _I),
        "ckd_gather_create": ([VP, _I, C.POINTER(VP)], _I), "ckd_gather_export": ([VP, VP], _I),
        "ckd_gather_open": ([VP, VP, C.POINTER(VP)], _I), "ckd_gather_destroy": ([VP], None),
        "ckd_gather_set_timeout_ms": ([VP, U], _I),
        "ckd_gather_acquire": ([VP, C.POINTER(VP)], _I), "ckd_gather_push": ([VP, VP, C.c_ulonglong], _I),
        "ckd_gather_pop": ([VP, C.c_ulonglong, _I, VP], _I), "ckd_gather_wait_pop": ([VP, C.c_ulonglong], _I),
        "ckd_gather_flush": ([VP], _I), "ckd_gather_status": ([VP], _I),
        "ckd_gather_checksums": ([VP, C.c_ulonglong, U, C.POINTER(C.c_ulonglong)], _I),
        "ckd_gather_peer_bytes": ([VP], C.c_ulonglong), "ckd_gather_slots": ([VP], _I),
        "ckd_frame_checksum": ([VP, VP, C.POINTER(C.c_ulonglong)], _I),
        "ckd_profile_begin": ([VP], _I),
        "ckd_profile_end": ([VP, C.POINTER(KernelStat), _I, C.POINTER(_I)], _I),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(L, name)  # raises AttributeError if the ABI symbol is not exported
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = L
    return L


ABI_SYMBOLS = None  # filled lazily by exported_symbols()


def geti(value):
    """Rocket::geti (rocket.h:27-29): int(roundf(float(value))) -- roundf rounds half away from zero"""
    v = float(np.float32(value))
    return int(np.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


def params_from_tracks(effect, track_value):
    """fills the POD struct of `effect` from a callable track_value(name) -> double (Rocket::getf/geti semantics)"""
    cls, names = TRACKS[effect]
    p = cls()
    kinds = dict(cls._fields_)
    for field, track in names.items():
        v = track_value(track)
        if kinds[field] is _I:
            setattr(p, field, geti(v))
        else:
            setattr(p, field, float(np.float32(v)))
    return p


class Context:
    """thin RAII wrapper around ckd_ctx*"""

    def __init__(self, res_x=1280, res_y=720, device=0, assets=None):
        self.L = load()
        h = C.c_void_p()
        self._check(self.L.ckd_create(C.byref(h), res_x, res_y, device))
        self.h = h
        self.res_x, self.res_y = res_x, res_y
        self.fx_x, self.fx_y = self.L.ckd_fxmap_res_x(h), self.L.ckd_fxmap_res_y(h)
        self._user = []
        if assets is not None:
            self.set_assets(assets)

    def _check(self, rc):
        if rc != 0:
            raise CkdError(f"ckd error {rc}: {self.L.ckd_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.ckd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- resources ----------------------------------------------------------------------------
    def set_assets(self, assets):
        for path, slots in IMAGE_SLOTS.items():
            arr = assets[path]
            bpp = 1 if arr.dtype == np.uint8 else 4
            for slot in (slots if isinstance(slots, tuple) else (slots,)):
                self._check(self.L.ckd_set_image(self.h, slot, arr.ctypes.data, arr.shape[1], arr.shape[0], bpp))

    def set_stream(self, cuda_stream):
        self._check(self.L.ckd_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._check(self.L.ckd_sync(self.h))

    def frame(self):
        return self.L.ckd_frame(self.h)

    def fxmap(self, i):
        return self.L.ckd_fxmap(self.h, i)

    def render_target(self, i):
        return self.L.ckd_render_target(self.h, i)

    def malloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.L.ckd_malloc(self.h, C.byref(p), nbytes))
        self._user.append(p.value)
        return p.value

    def free(self, ptr):
        self._check(self.L.ckd_free(self.h, C.c_void_p(ptr)))
        self._user.remove(ptr)

    def upload(self, d_ptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self.L.ckd_upload(self.h, C.c_void_p(d_ptr), arr.ctypes.data, arr.nbytes))
        self.sync()  # the host array may be a temporary

    def download(self, d_ptr, shape, dtype=np.uint32):
        out = np.empty(shape, dtype=dtype)
        self._check(self.L.ckd_download(self.h, out.ctypes.data, C.c_void_p(d_ptr), out.nbytes))
        self.sync()
        return out

    def to_device(self, arr, pad_elems=64):
        """copies a host array into a fresh device buffer with `pad_elems` zeroed elements of slack after it"""
        arr = np.ascontiguousarray(arr)
        nbytes = arr.nbytes + pad_elems * arr.itemsize
        d = self.malloc(nbytes)
        padded = np.zeros(arr.size + pad_elems, dtype=arr.dtype)
        padded[:arr.size] = arr.ravel()
        self.upload(d, padded)
        return d

    def rsqrt_table(self):
        log2 = C.c_int()
        n = C.c_size_t()
        self._check(self.L.ckd_get_rsqrt_table(self.h, None, 0, C.byref(log2), C.byref(n)))
        tab = np.zeros(n.value, dtype=np.uint32)
        self._check(self.L.ckd_get_rsqrt_table(self.h, tab.ctypes.data_as(_U32P), tab.size, C.byref(log2), C.byref(n)))
        return tab, log2.value

    def set_rsqrt_table(self, tab, log2_bin):
        tab = np.ascontiguousarray(tab, dtype=np.uint32)
        self._check(self.L.ckd_set_rsqrt_table(self.h, tab.ctypes.data_as(_U32P), log2_bin))

    def polar_maps(self):
        m = np.zeros((self.res_y, self.res_x, 2), dtype=np.int32)
        inv = np.zeros_like(m)
        self._check(self.L.ckd_get_polar_maps(self.h, m.ctypes.data, inv.ctypes.data))
        return m, inv

    def launch_count(self):
        return int(self.L.ckd_launch_count(self.h))

    # -- effects ------------------------------------------------------------------------------
    def draw(self, effect, params, time, d_dest=None, close=None):
        d_dest = d_dest or self.frame()
        L, h, t = self.L, self.h, C.c_float(time)
        if effect == "spikey":
            rc = L.ckd_spikey_draw(h, C.byref(params), t, int(bool(close)), C.c_void_p(d_dest))
        else:
            rc = getattr(L, f"ckd_{effect}_draw")(h, C.byref(params), t, C.c_void_p(d_dest))
        self._check(rc)

    def ball_beam_tail(self, d_rows, row_pixels, rows, first_remainder, beam_color, beam_alpha_min, raw_steps=False):
        """the beam tail of vball_ray_beams alone (ball.cpp:168-203), one row per tail length; raw_steps: the float bits of curStep"""
        self._check(self.L.ckd_ball_beam_tail(self.h, C.c_void_p(d_rows), row_pixels, rows, first_remainder, beam_color, C.c_float(beam_alpha_min), int(bool(raw_steps))))

    def read_frame(self, d_ptr=None):
        return self.download(d_ptr or self.frame(), (self.res_y, self.res_x))

    # -- post ops -----------------------------------------------------------------------------
    def fx_blit_2x2(self, d_dest, d_src):
        self._check(self.L.ckd_fx_blit_2x2(self.h, C.c_void_p(d_dest), C.c_void_p(d_src)))

    def polar_blit(self, d_dest, d_src, inverse=False, alpha=False):
        fn = self.L.ckd_polar_blit_a if alpha else self.L.ckd_polar_blit
        self._check(fn(self.h, C.c_void_p(d_dest), C.c_void_p(d_src), int(inverse)))

    def polar_blit_2x2(self, d_dest, d_src, inverse=False):
        self._check(self.L.ckd_polar_blit_2x2(self.h, C.c_void_p(d_dest), C.c_void_p(d_src), int(inverse)))

    def blend_chain(self, d_dest, steps, n):
        """steps: [(op name, d_src or None, f_param, u_param)] applied per pixel in one pass (ckd_blend_chain)"""
        arr = (BlendStep * len(steps))()
        for a, (op, d_src, f, u) in zip(arr, steps):
            a.op, a.d_src, a.f_param, a.u_param = BLEND_OPS[op], d_src, f, u
        self._check(self.L.ckd_blend_chain(self.h, C.c_void_p(d_dest), arr, len(steps), n))

    def old_blur(self, kind, d_dest, d_src, w, h, strength):
        fn = {"h": self.L.ckd_old_blur_h, "v": self.L.ckd_old_blur_v, "hv": self.L.ckd_old_blur}[kind]
        self._check(fn(self.h, C.c_void_p(d_dest), C.c_void_p(d_src), w, h, C.c_float(strength)))

    def new_blur(self, kind, d_dest, d_src, w, h, strength, gain, passes):
        fn = {"h": self.L.ckd_new_blur_h, "v": self.L.ckd_new_blur_v, "hv": self.L.ckd_new_blur}[kind]
        self._check(fn(self.h, C.c_void_p(d_dest), C.c_void_p(d_src), w, h, C.c_float(strength), C.c_float(gain), passes))

    def blend(self, op, d_dest, d_src, n, fparam=0.0, uparam=0):
        self._check(self.L.ckd_blend(self.h, BLEND_OPS[op], C.c_void_p(d_dest), C.c_void_p(d_src), n, C.c_float(fparam), C.c_uint(uparam)))

    def blit(self, op, d_dest, d_src, dest_res_x, src_res_x, y_res, alpha=1.0):
        self._check(self.L.ckd_blit(self.h, BLIT_OPS[op], C.c_void_p(d_dest), C.c_void_p(d_src), dest_res_x, src_res_x, y_res, C.c_float(alpha)))

    def memset32(self, d_dest, value, n):
        self._check(self.L.ckd_memset32(self.h, C.c_void_p(d_dest), C.c_uint32(value), n))

    def tape_warp(self, d_dest, d_src, w, h, strength, speed):
        self._check(self.L.ckd_tape_warp(self.h, C.c_void_p(d_dest), C.c_void_p(d_src), w, h, C.c_float(strength), C.c_float(speed)))

    def profile_begin(self):
        self._check(self.L.ckd_profile_begin(self.h))

    def profile_end(self):
        """-> {kernel name: {"launches", "total_ms", "algo_bytes"}} measured with CUDA events around every launch"""
        stats = (KernelStat * 64)()
        count = C.c_int()
        self._check(self.L.ckd_profile_end(self.h, stats, 64, C.byref(count)))
        return {stats[i].name.decode(): {"launches": int(stats[i].launches), "total_ms": float(stats[i].total_ms), "algo_bytes": float(stats[i].algo_bytes)}
                for i in range(count.value)}

    # -- helpers / frame gather -----------------------------------------------------------------
    def fastcos(self, x, sine=False):
        """fastcosf / fastsinf (fast-cosine.h:17-53) of a float64 array, evaluated on the device"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        d_x = self.to_device(x, pad_elems=0)
        d_out = self.malloc(max(4, x.size * 4))
        self._check(self.L.ckd_fastcos(self.h, C.c_void_p(d_out), C.c_void_p(d_x), x.size, int(bool(sine))))
        out = self.download(d_out, x.shape, dtype=np.float32)
        self.free(d_x)
        self.free(d_out)
        return out

    def fast_cos_table(self):
        tab = np.zeros(1025, dtype=np.float64)
        self._check(self.L.ckd_get_fast_cos_table(self.h, tab.ctypes.data))
        return tab

    def set_frame_independent(self, enabled):
        self._check(self.L.ckd_set_frame_independent(self.h, int(bool(enabled))))

    def frame_checksum(self, d_ptr=None):
        """sum_i pixel[i]*(2i+1) mod 2^64 of a device frame (the checksum the gather's collector computes)"""
        out = C.c_ulonglong()
        self._check(self.L.ckd_frame_checksum(self.h, C.c_void_p(d_ptr or self.frame()), C.byref(out)))
        return int(out.value)

    def malloc_host(self, nbytes):
        p = C.c_void_p()
        self._check(self.L.ckd_malloc_host(C.byref(p), nbytes))
        return p.value

    def free_host(self, ptr):
        self._check(self.L.ckd_free_host(C.c_void_p(ptr)))

    def timer_start(self):
        self._check(self.L.ckd_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float()
        self._check(self.L.ckd_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value


GATHER_CHECKSUM, GATHER_TO_HOST = 1, 2
GATHER_HANDLE_BYTES = 128


def frame_checksum_host(frame):
    """the gather's frame checksum computed on the host: sum_i pixel[i]*(2i+1) mod 2^64"""
    px = np.ascontiguousarray(frame).reshape(-1).astype(np.uint64)
    w = np.arange(px.size, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    with np.errstate(over="ignore"):
        return int((px * w).sum(dtype=np.uint64))


class Gather:
    """ckd_gather: the slot ring in the collector's HBM that the producers fill with peer copies (include/ckd.h).
    Gather(ctx, slots=8) creates it (collector); Gather(ctx, handle=bytes) maps it from another process (producer)."""

    def __init__(self, ctx, slots=8, handle=None):
        self.L, self.ctx = ctx.L, ctx
        g = C.c_void_p()
        if handle is None:
            ctx._check(self.L.ckd_gather_create(ctx.h, slots, C.byref(g)))
        else:
            buf = C.create_string_buffer(bytes(handle), GATHER_HANDLE_BYTES)
            ctx._check(self.L.ckd_gather_open(ctx.h, buf, C.byref(g)))
        self.g = g

    def export(self):
        buf = C.create_string_buffer(GATHER_HANDLE_BYTES)
        self.ctx._check(self.L.ckd_gather_export(self.g, buf))
        return bytes(buf.raw)

    def set_timeout_ms(self, ms):
        self.ctx._check(self.L.ckd_gather_set_timeout_ms(self.g, int(ms)))

    def acquire(self):
        p = C.c_void_p()
        self.ctx._check(self.L.ckd_gather_acquire(self.g, C.byref(p)))
        return p.value

    def push(self, seq, d_frame=None):
        self.ctx._check(self.L.ckd_gather_push(self.g, C.c_void_p(d_frame), seq))

    def pop(self, seq, mode=0, h_dest=None):
        self.ctx._check(self.L.ckd_gather_pop(self.g, seq, mode, C.c_void_p(h_dest)))

    def wait_pop(self, seq):
        self.ctx._check(self.L.ckd_gather_wait_pop(self.g, seq))

    def flush(self):
        self.ctx._check(self.L.ckd_gather_flush(self.g))

    def status(self):
        self.ctx._check(self.L.ckd_gather_status(self.g))

    def checksums(self, first_seq, count):
        out = (C.c_ulonglong * count)()
        self.ctx._check(self.L.ckd_gather_checksums(self.g, first_seq, count, out))
        return [int(v) for v in out]

    def peer_bytes(self):
        return int(self.L.ckd_gather_peer_bytes(self.g))

    def close(self):
        if self.g:
            self.L.ckd_gather_destroy(self.g)
            self.g = None
