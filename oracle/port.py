"""oracle/port.py -- TEST INFRASTRUCTURE: ctypes driver of the plain-C restatement (oracle/ckd_oracle.c).

Re-creates each reference X_Draw on top of the C primitives (render map -> Fx_Blit_2x2 -> optional blur / blends ...),
following the same reference lines the C file cites.  Pinned against tests/golden by tests/test_oracle_port.py.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libckd_oracle.so")

_U32P = C.POINTER(C.c_uint32)
_f32 = np.float32


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _farr(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def _clampf(mn, mx, v):
    return max(mn, min(mx, v))


class Port:
    def __init__(self, res_x, res_y, rsqrt_table, assets):
        if not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "ckd_oracle.c")):
            build()
        L = self.L = C.CDLL(LIB)
        self.res_x, self.res_y = res_x, res_y
        self.fx_x, self.fx_y = res_x // 2 + 4, res_y // 2 + 4
        self.assets = assets
        self._rsqrt = np.ascontiguousarray(rsqrt_table, dtype=np.uint32)
        L.orc_init(res_x, res_y, _p(self._rsqrt))
        L.orc_lutcosf.restype = C.c_float
        L.orc_lutcosf.argtypes = [C.c_float]
        L.orc_rsqrt.restype = C.c_float
        L.orc_rsqrt.argtypes = [C.c_float]
        L.orc_log_ps.restype = C.c_float
        L.orc_log_ps.argtypes = [C.c_float]
        L.orc_exp_ps.restype = C.c_float
        L.orc_exp_ps.argtypes = [C.c_float]
        L.orc_gamma_pixel.restype = C.c_uint32
        L.orc_gamma_pixel.argtypes = [C.POINTER(C.c_float), C.c_float]
        L.orc_box_blur_scale.restype = C.c_float
        L.orc_box_blur_scale.argtypes = [C.c_float]
        L.orc_cos_lut.restype = C.POINTER(C.c_float)
        VP, F, U, I = C.c_void_p, C.c_float, C.c_uint, C.c_int
        L.orc_fx_blit_2x2.argtypes = [VP, VP]
        L.orc_polar_maps.argtypes = [VP, VP]
        L.orc_polar_blit.argtypes = [VP, VP, VP, I]
        L.orc_polar_maps_2x2.argtypes = [VP, VP]
        L.orc_polar_blit_2x2.argtypes = [VP, VP, VP]
        for n in ("orc_old_blur_h", "orc_old_blur_v", "orc_old_blur"):
            getattr(L, n).argtypes = [VP, VP, U, U, F]
        L.orc_new_blur.argtypes = [I, VP, VP, C.c_size_t, U, U, F, F, U]
        L.orc_blend.argtypes = [I, VP, VP, U, F, U]
        L.orc_blit.argtypes = [I, VP, VP, U, U, U, U, F]
        L.orc_tape_warp.argtypes = [VP, VP, U, U, F, F]
        L.orc_plasma_map.argtypes = [VP, F, F, F, F, F]
        L.orc_nautilus_map.argtypes = [VP, F, F, F, F, F]
        L.orc_spikey_map.argtypes = [I, VP, F, C.POINTER(F)]
        L.orc_sinuses_map.argtypes = [VP, F, F, F, F, F, F, F, F]
        L.orc_laura_map.argtypes = [VP, F, F, F, F, F, F, F]
        L.orc_tunnel_map.argtypes = [VP, VP, VP, VP, F, C.POINTER(F)]
        L.orc_landscape.argtypes = [VP, VP, VP, VP, F, F]
        L.orc_tunnelscape_rt.argtypes = [VP, VP, VP, VP, F, F, F, F]
        L.orc_ball_rt.argtypes = [VP, VP, VP, VP, C.POINTER(I), C.POINTER(F)]
        L.orc_twister_rt.argtypes = [VP, VP, VP, F, F, F]
        L.orc_cspan16.argtypes = [VP, I, U, U, C.c_uint32, C.c_uint32]
        self._maps = None
        self.render_target0 = np.zeros((res_y, res_x), dtype=np.uint32)

    # -- buffers ------------------------------------------------------------------------------------
    def buf(self, n, pad=None):
        pad = 4 * self.res_x if pad is None else pad
        return np.zeros(n + pad, dtype=np.uint32)[:n]

    def frame(self):
        return self.buf(self.res_x * self.res_y).reshape(self.res_y, self.res_x)

    def fxmap(self):
        return self.buf(self.fx_x * self.fx_y).reshape(self.fx_y, self.fx_x)

    def polar_maps(self):
        if self._maps is None:
            m = np.zeros((self.res_y, self.res_x, 2), dtype=np.int32)
            inv = np.zeros_like(m)
            self.L.orc_polar_maps(_p(m), _p(inv))
            self._maps = (m, inv)
        return self._maps

    # -- post ops (same names as the reference) -----------------------------------------------------------
    def fx_blit_2x2(self, dst, src):
        self.L.orc_fx_blit_2x2(_p(dst), _p(src))

    def polar_blit(self, dst, src, inverse=False, alpha=False):
        m, inv = self.polar_maps()
        self.L.orc_polar_blit(_p(dst), _p(src), _p(inv if inverse else m), int(alpha))

    def polar_blit_2x2(self, dst, src, inverse=False):
        if getattr(self, "_maps2x2", None) is None:
            m = np.zeros((self.fx_y, self.fx_x, 2), dtype=np.int32)
            inv = np.zeros_like(m)
            self.L.orc_polar_maps_2x2(_p(m), _p(inv))
            self._maps2x2 = (m, inv)
        self.L.orc_polar_blit_2x2(_p(dst), _p(src), _p(self._maps2x2[1 if inverse else 0]))

    def old_blur(self, kind, dst, src, w, h, strength):
        fn = {"h": self.L.orc_old_blur_h, "v": self.L.orc_old_blur_v, "hv": self.L.orc_old_blur}[kind]
        fn(_p(dst), _p(src), w, h, C.c_float(strength))

    def new_blur(self, kind, dst, src, w, h, strength, gain, passes, src_elems=None):
        self.L.orc_new_blur({"h": 0, "v": 1, "hv": 2}[kind], _p(dst), _p(src), src_elems if src_elems is not None else w * h, w, h,
                            C.c_float(strength), C.c_float(gain), passes)

    def blend(self, op, dst, src, fparam=0.0, uparam=0, n=None):
        from oracle.ref import BLEND_OPS
        self.L.orc_blend(BLEND_OPS[op], _p(dst), _p(src) if src is not None else None, dst.size if n is None else n, C.c_float(fparam), C.c_uint(uparam))

    def blit(self, op, dst, src, dest_res_x, src_res_x, y_res, alpha=1.0):
        from oracle.ref import BLIT_OPS
        self.L.orc_blit(BLIT_OPS[op], _p(dst), _p(src), dest_res_x, src_res_x, src_res_x, y_res, C.c_float(alpha))

    def mix_src_s(self, dst, src, dest_res_x, dest_res_y, src_stride):
        self.L.orc_blit(0, _p(dst), _p(src), dest_res_x, src_stride, dest_res_x, dest_res_y, C.c_float(1.0))

    def tape_warp(self, dst, src, w, h, strength, speed):
        self.L.orc_tape_warp(_p(dst), _p(src), w, h, C.c_float(strength), C.c_float(speed))

    def box_blur_scale(self, s):
        return float(self.L.orc_box_blur_scale(C.c_float(s)))

    # -- X_Draw equivalents ---------------------------------------------------------------------------------
    def draw(self, effect, p, time, close=None):
        """p: dict of the evaluated parameter struct (same field names as include/ckd.h); returns the finished frame"""
        L, A = self.L, self.assets
        W, H = self.res_x, self.res_y
        t = C.c_float(time)
        dest = self.frame()
        fx0 = self.fxmap()
        f = lambda k: C.c_float(p[k])  # noqa: E731

        if effect == "plasma":  # Plasma_Draw, shadertoy.cpp:276-280
            L.orc_plasma_map(_p(fx0), t, f("speed"), f("hue"), f("gamma"), f("desaturation"))
            self.fx_blit_2x2(dest, fx0)
        elif effect == "nautilus":  # Nautilus_Draw, shadertoy.cpp:395-407
            L.orc_nautilus_map(_p(fx0), t, f("roll"), f("hue"), f("speed"), f("desaturation"))
            self.fx_blit_2x2(dest, fx0)
            blur = self.box_blur_scale(p["blur"])
            if blur != 0.0:
                self.old_blur("hv", dest, dest, W, H, blur)
        elif effect == "sinuses":
            L.orc_sinuses_map(_p(fx0), t, f("specular"), f("roll"), f("speed"), f("offs_x"), f("gamma"), f("hue"), f("desaturation"))
            self.fx_blit_2x2(dest, fx0)
        elif effect == "laura":
            L.orc_laura_map(_p(fx0), t, f("speed"), f("yaw"), f("pitch"), f("roll"), f("hue"), f("saturate"))
            self.fx_blit_2x2(dest, fx0)
        elif effect == "spikey":  # Spikey_Draw, shadertoy.cpp:661-733
            def pack(x, y, z):
                return _farr([p["speed"], p["roll"], p["specular"], p["desaturation"], p["hue"], p["gamma"], x, y, z, p["close_z_scale"],
                              p["close_normal_grain"], p["close_scale"], p["close_aspect_mul"], 1.0 + p["warmup"]])
            if close:
                L.orc_spikey_map(0, _p(fx0), t, pack(p["close_x"], p["close_y"], p["close_z"]))
                opacity = float(_f32(_clampf(0.0, 1.0, p["mix_blur_opacity"])))
                if opacity > 0.0:
                    n = self.fx_x * self.fx_y
                    mb_map = _clampf(0.0, 1.0, p["mix_blur_map"])
                    mb_blur = _clampf(0.0, 100.0, p["mix_blur"])
                    mb_map_blur = _clampf(0.0, 100.0, p["mix_map_blur"])
                    m0 = A["assets/shadertoy/close-up-blur-map-1.png"].ravel()
                    m1 = A["assets/shadertoy/close-up-blur-map-2.png"].ravel()
                    blur_map = self.buf(n)
                    if mb_map == 0.0:
                        blur_map[:] = m0
                    elif mb_map == 1.0:
                        blur_map[:] = m1
                    else:
                        blur_map[:] = m0
                        src = self.buf(n); src[:] = m1
                        self.blend("Mix32", blur_map, src, uparam=int(_f32(mb_map) * _f32(255.0)) & 0xff)
                    fx1 = self.buf(n); fx1[:] = fx0.ravel()
                    if mb_map_blur >= 1.0:
                        self.old_blur("hv", blur_map, blur_map, self.fx_x, self.fx_y, self.box_blur_scale(mb_map_blur))
                    if mb_blur >= 1.0:
                        self.old_blur("hv", fx1, fx1, self.fx_x, self.fx_y, self.box_blur_scale(mb_blur))
                    self.blend("SoftLight32AA", fx1, blur_map, fparam=float(np.tanh(_f32(mb_blur) + _f32(opacity), dtype=np.float32)))
                    flat = fx0.reshape(-1)
                    self.blend("Overlay32A", flat, fx1)
                self.fx_blit_2x2(dest, fx0)
            elif p["warmup"] == 0.0:
                L.orc_spikey_map(1, _p(fx0), t, pack(p["dist_x"], p["dist_y"], p["dist_z"]))
                self.fx_blit_2x2(dest, fx0)
            else:
                L.orc_spikey_map(2, _p(fx0), t, pack(0, 0, 0))
                flat = fx0.reshape(-1)
                self.old_blur("h", flat, flat, self.fx_x, self.fx_y, self.box_blur_scale(float(_f32(1.0) + _f32(p["warmup"]))))
                self.fx_blit_2x2(dest, fx0)
        elif effect == "tunnel":  # Tunnel_Draw, shadertoy.cpp:840-861
            fx1 = self.fxmap()
            pp = _farr([p["boxy"], p["flower_scale"], p["flower_freq"], p["flower_phase"], p["speed"], p["roll"], p["pitch"], p["radius"],
                        p["mul_u"], p["mul_v"], p["fog1"], p["fog2"]])
            L.orc_tunnel_map(_p(fx0), _p(fx1), _p(A["assets/shadertoy/nytrik-hextexture.png"]), _p(A["assets/shadertoy/nytrik-hextexture-fx.png"]), t, pp)
            if p["lit_tiles"] != 0:
                lit_blur = _clampf(0.0, 100.0, p["lit_blur"])
                flat1 = fx1.reshape(-1)
                if lit_blur >= 1.0:
                    self.old_blur("hv", flat1, flat1, self.fx_x, self.fx_y, self.box_blur_scale(lit_blur))
                self.blend("Add32", fx0.reshape(-1), flat1)
            self.fx_blit_2x2(dest, fx0)
        elif effect == "landscape":  # Landscape_Draw, landscape.cpp:228-243
            warp = p["warp_strength"] != 0.0
            target = self.frame() if warp else dest
            L.orc_landscape(_p(target), _p(A["assets/scape/D17.png"]), _p(A["assets/scape/C17W-edit.png"]), _p(A["assets/scape/foggradient.jpg"]),
                            f("forward"), f("tilt"))
            if warp:
                self.tape_warp(dest, target, W, H, p["warp_speed"], p["warp_strength"])
        elif effect == "tunnelscape":  # Tunnelscape_Draw, tunnelscape.cpp:168-186
            rt = self.frame()
            L.orc_tunnelscape_rt(_p(rt), _p(A["assets/scape/tscape-D7-edit.png"]), _p(A["assets/scape/tscape-C7W-edit.png"]), _p(A["assets/scape/foggradient.jpg"]),
                                 t, f("step_u"), f("step_v"), f("speed"))
            self.polar_blit(dest, rt, inverse=True)
            if p["blur"] != 0.0:
                s = self.box_blur_scale(p["blur"])
                self.old_blur("hv", dest, dest, W, H, s)
                self.old_blur("hv", dest, dest, W, H, s)
        elif effect == "ball":  # Ball_Draw, ball.cpp:452-514
            has_beams = p["has_beams"] != 0
            paths = ["assets/ball/hmap_1_1k.jpg", "assets/ball/hmap_4_1k.jpg", "assets/ball/hmap_2_1k.jpg", "assets/ball/hmap_3_1k.jpg", "assets/ball/hmap_5_1k.jpg"]
            base = max(1, min(4, p["base_shape_index"]))
            mix = np.zeros(1024 * 1024 + 4096, dtype=np.uint8)[:1024 * 1024]
            mix[:] = A[paths[base]].ravel()
            spikes = p["spikes"] & 0xff
            if spikes != 0:
                src = np.zeros(1024 * 1024 + 4096, dtype=np.uint8)[:1024 * 1024]
                src[:] = A[paths[0]].ravel()
                self.blend("Mix32", mix.view(np.uint32), src.view(np.uint32), uparam=spikes)
            if has_beams:
                aux = self.buf(1024 * 1024)
                for key, path in (("beams1", "assets/ball/beammap_1k_1.jpg"), ("beams2", "assets/ball/beammap_1k_2.jpg"), ("beams3", "assets/ball/beammap_1k_3-2.jpg")):
                    a = _clampf(0.0, 1.0, p[key])
                    if a > 0.0:
                        self.blit("BlitAdd32A", aux, np.ascontiguousarray(A[path]).ravel(), 1024, 1024, 1024, a)
                color = A["assets/ball/colormap_1k.jpg"]
            else:
                aux = np.ascontiguousarray(A["assets/ball/envmap3_1k.jpg"]).ravel()
                color = A["assets/ball/colormap_2_1k.jpg"]
            rt = self.frame()
            rt[:] = self.render_target0
            ip = (C.c_int * 4)(max(1, min(1024, p["ray_length"])), max(0, min(255, p["beam_atten"])), max(0, min(255, p["low_beams"])), int(has_beams))
            fp = _farr([_clampf(1.0, 1920.0, p["radius"]), _clampf(0.0, 255.0, p["beam_alpha_min"]), float(_f32(time) * _f32(p["speed"])),
                        p["rotate_offs_x"], p["rotate_offs_y"]])
            L.orc_ball_rt(_p(rt), _p(mix), _p(color), _p(aux), ip, fp)
            blur = self.box_blur_scale(p["blur"])
            if blur != 0.0:
                self.old_blur("h", rt, rt, W, H, blur)
            dest[:] = A["assets/ball/nytrik-background_1280x720.png" if has_beams else "assets/ball/nytrik-background-2-1280x720.png"]
            self.polar_blit(dest, rt, inverse=False, alpha=True)
            if has_beams:
                halo = self.frame(); halo[:] = A["assets/ball/halo.png"]
                self.blend("SoftLight32A", dest.reshape(-1), halo.reshape(-1))
        elif effect == "twister":  # Twister_Draw, torus-twister.cpp:166-188
            rt = self.frame()
            L.orc_twister_rt(_p(rt), _p(A["assets/twister/hmap_2_1k.jpg"]), _p(A["assets/twister/colormap_1k.jpg"]), t, f("speed"), f("shear_speed"))
            if p["blur"] != 0.0:
                self.old_blur("h", rt, rt, W, H, self.box_blur_scale(p["blur"]))
            dest[:] = A["assets/twister/nytrik-background_1280x720.png"]
            self.polar_blit(dest, rt, inverse=False, alpha=True)
        else:
            raise ValueError(effect)
        return dest


# ---------------------------------------------------------------------------------------------------------------------
# the compositor: Demo_Draw, code/demo.cpp:469-1023 (BloodBlend/CreditBlend 393-467, FadeFlash 383-390) on top of the
# primitives above.  T: oracle.rocket.Tracks positioned at the frame's time (Demo_Draw runs Rocket::Boost itself).
# ---------------------------------------------------------------------------------------------------------------------

_libm = C.CDLL("libm.so.6")
for _n in ("powf", "fmodf"):
    getattr(_libm, _n).restype = C.c_float
    getattr(_libm, _n).argtypes = [C.c_float, C.c_float]
_libm.sinf.restype = C.c_float
_libm.sinf.argtypes = [C.c_float]

K_PI = _f32(3.1415926535897932384626433832795)
K_2PI = _f32(2.0) * K_PI
K_GOLDEN_RATIO = _f32(1.61803398875)
K_GOLDEN_ANGLE = _f32(2.39996)


def _sat(v):
    return _f32(max(0.0, min(1.0, float(v))))


def _u8(v):
    """float -> uint8_t as gcc compiles it on x86-64 (cvttss2si, low byte)"""
    v = float(v)
    return (int(v) & 0xFF) if -2147483648.0 <= v < 2147483648.0 else 0


def _effect_params(effect, T):
    from cookiedough_b200 import capi
    cls, names = capi.TRACKS[effect]
    types = dict(cls._fields_)
    p = {name: 0 for name, _ in cls._fields_}
    for field, track in names.items():
        p[field] = T.geti(track) if types[field] is C.c_int else T.getf(track)
    return p


def _demo_draw(self, time_s, T):
    A, W, H = self.assets, self.res_x, self.res_y
    N = W * H
    T.set_time(time_s)
    if T.get("demo:quit") != 0.0:
        return None  # demo is over (code/rocket.cpp:78-79)
    timer = float(_f32(time_s))

    def layer(path):
        a = self.buf(A[path].size)
        a[:] = np.ascontiguousarray(A[path]).ravel()
        return a

    def effect(name, ckd_effect, close=None):
        return self.draw(ckd_effect, _effect_params(ckd_effect, T), timer, close=close).reshape(-1)

    def fade_flash(d, to_black, to_white):
        if to_white > 0.0:
            self.blend("Fade32", d, None, uparam=(_u8(_f32(to_white) * _f32(255.0)) << 24) | 0xFFFFFF, n=N)
        if to_black > 0.0:
            self.blend("Fade32", d, None, uparam=(_u8(_f32(to_black) * _f32(255.0)) << 24), n=N)

    def full(op, d, path, f=0.0):
        self.blend(op, d, layer(path), fparam=f, n=N)

    def logo_blend(blend, paths, width, height):
        last = len(paths) - 1
        factor = _libm.fmodf(C.c_float(blend), C.c_float(1.0))
        i_factor = _u8(_f32(255.0) * _f32(factor))
        if blend >= float(last):
            return layer(paths[last])
        target = self.buf(N)
        for i in range(last):
            if float(i) <= blend < float(i + 1):
                src = layer(paths[i])
                target[:src.size] = src
                self.blend("Mix32", target, layer(paths[i + 1]), uparam=i_factor, n=width * height)
                break
        return target

    fade_black, fade_white = T.getf("demo:FadeToBlack"), T.getf("demo:FadeToWhite")
    part = T.geti("demo:Effect")

    if part == 1:
        d = effect("twister", "twister")
        fade_flash(d, fade_black, fade_white)
        full("SoftLight32A", d, "assets/closeup/Vignette_CoolFilmLook.png")
        full("MulSrc32A", d, "assets/demo/tpb-06-dirty-vignette-1280x720.png")
    elif part == 2:
        d = effect("landscape", "landscape")
        fade_flash(d, float(_sat(T.getf("demo:ScapeFade"))), 0.0)
        if T.geti("shootingStar:Enabled") == 1:
            x, y, alpha = T.geti("shootingStar:X"), T.geti("shootingStar:Y"), _f32(T.getf("shootingStar:A"))
            lenz = layer("assets/shooting/Lenz.png")
            self.blit("BlitAdd32A", d[y * W + x:], lenz, W, 64, 64, float(alpha))
            trail = T.geti("shootingStar:Trail")
            if trail > 0:
                alpha_step = _f32(alpha / _f32(trail))
                for _ in range(trail):
                    x += 4; y -= 1; alpha = _f32(alpha - alpha_step)
                    self.blit("BlitAdd32A", d[y * W + x:], lenz, W, 64, 64, float(alpha))
        overlay = _sat(T.getf("demo:ScapeOverlay"))
        if overlay != 0.0:
            self.blit("BlitAdd32A", d, layer("assets/demo/nytrik-god-layer-720p.png"), W, W, H, float(overlay))
        rev = _sat(T.getf("demo:ScapeRev"))
        if rev != 0.0:
            rt0 = self.buf(N)
            logo = layer("assets/scape/revision-logo_white.png")
            if rev < _f32(0.314):
                c4 = _f32(_f32(2.0) * K_PI) / _f32(3.0)
                if rev == 0.0:
                    ease_a = _f32(0.0)
                elif rev == 1.0:
                    ease_a = _f32(1.0)
                else:
                    ease_a = _f32(_f32(_libm.powf(C.c_float(2.0), C.c_float(_f32(-10.0) * rev))) * _f32(_libm.sinf(C.c_float(_f32(_f32(rev * _f32(10.0)) - _f32(0.75)) * c4))) + _f32(1.0))
                c1 = _f32(1.70158); c3 = _f32(c1 + _f32(1.0))
                ease_b = _f32(_f32(_f32(_f32(c3 * rev) * rev) * rev) - _f32(_f32(c1 * rev) * rev))
                self.tape_warp(rt0, logo, W, H, float(_f32(ease_a * K_GOLDEN_ANGLE)), float(_f32(ease_b * K_GOLDEN_RATIO)))
            else:
                self.old_blur("hv", rt0, logo, W, H, self.box_blur_scale(float(_f32(_f32(rev - _f32(0.314)) * K_2PI))))
            self.blit("BlitSrc32A", d, rt0, W, W, H, float(rev))
        fade_flash(d, fade_black, fade_white)
        full("SoftLight32A", d, "assets/closeup/Vignette_CoolFilmLook.png")
    elif part == 3:
        d = effect("ball", "ball")
        beams = T.geti("ball:HasBeams") != 0
        if not beams:
            full("MulSrc32", d, "assets/greetings/Vignette_CoolFilmLook.png")
        else:
            full("SoftLight32", d, "assets/ball/Vignette_Sparta300.png")
        fade_flash(d, fade_black, fade_white)
        if beams:
            full("MulSrc32A", d, "assets/demo/tpb-06-dirty-vignette-1280x720.png")
    elif part == 4:
        d = effect("tunnelscape", "tunnelscape")
        full("Sub32", d, "assets/tunnels/Vignette_Layer02_inverted.png")
        full("MixSrc32", d, "assets/tunnels/nytrik-TheYearWas_Overlay_LensDirt.png")
        show = _clampf(0.0, 3.0, T.getf("demo:Show1995"))
        if show > 0.0:
            self.blend("MixOver32", d, logo_blend(show, [f"assets/tunnels/layer 1995_{i}.png" for i in range(1, 5)], W, H), n=N)
        full("Overlay32", d, "assets/tunnels/Vignette_CoolFilmLook.png")
    elif part == 5:
        d = effect("plasma", "plasma")
        i_logo = max(0, min(4, T.geti("demo:CreditLogo")))
        if i_logo != 0:
            blend = _clampf(0.0, 4.0, T.getf("demo:CreditAnimBlend"))
            stem = {1: "assets/credits/animplek/animplek{}.png", 2: "assets/credits/comatron_anim/comatron_{}.png",
                    3: "assets/credits/jade&nytrik/jade&nytrik{}.png", 4: "assets/credits/animhot0/animhot{}.png"}[i_logo]
            paths = [stem.format(i + 1 if i_logo == 2 else i) for i in range(5)]
            cur = logo_blend(blend, paths, 1280, 568)
            blur_h = T.getf("demo:CreditLogoBlurH")
            if blur_h != 0.0:
                rt0 = self.buf(N)
                self.old_blur("h", rt0, cur, 1280, 568, self.box_blur_scale(blur_h))
                cur = rt0
            blur_v = T.getf("demo:CreditLogoBlurV")
            if blur_v != 0:
                rt0 = cur if blur_h != 0.0 else self.buf(N)
                self.old_blur("v", rt0, cur, 1280, 568, self.box_blur_scale(blur_v))
                cur = rt0
            self.blit("BlitSrc32A", d[((H - 568) >> 1) * W:], cur, W, 1280, 568, _clampf(0.0, 1.0, T.getf("demo:CreditLogoAlpha")))
    elif part == 6:
        d = effect("nautilus", "nautilus")
        full("SoftLight32", d, "assets/nautilus/Vignette.png")
        full("SoftLight32", d, "assets/nautilus/GlassDirt_Distorted2.png")
        fade_flash(d, fade_black, 0.0)
        first = T.geti("demo:Cousteau") == 0
        cousteau = layer("assets/nautilus/JacquesCousteau1_Silhouette.png" if first else "assets/nautilus/JacquesCousteau_Silhouette2.png")
        full("Overlay32A", d, "assets/nautilus/JacquesCousteau1_Silhouette_RimMask.png" if first else "assets/nautilus/JacquesCousteau_Silhouette2_RimMask.png")
        h_blur = T.getf("demo:CousteauHorzBlur")
        if h_blur != 0.0:
            rt0 = self.buf(N)
            self.old_blur("h", rt0, cousteau, W, H, self.box_blur_scale(h_blur))
            cousteau = rt0
        self.blend("MixSrc32", d, cousteau, n=N)
        fade_flash(d, 0.0, fade_white)
        full("MixSrc32", d, "assets/nautilus/JacquesCousteau_Text.png")
    elif part == 7:
        d = effect("spikey_close", "spikey", close=True)
        dirt = T.geti("demo:Dirt")
        if dirt != 1:
            full("MulSrc32", d, "assets/spikeball/Vignette_CoolFilmLook.png")
        if dirt == 1:
            raker = T.getf("closeSpike:Moonraker")
            raker_text = _clampf(0.0, 2.0, T.getf("closeSpike:MoonrakerText"))
            if raker > 0.0:
                full("MulSrc32", d, "assets/closeup/VignetteForRaker.png")
                full("SoftLight32AA", d, "assets/closeup/raker-LensDirt5_invert.png", raker)
                text = layer("assets/closeup/raker_textSmall.png")
                if 0.0 < raker_text < 1.0:
                    rt2 = self.buf(N)
                    self.blit("BlitSrc32", rt2[(H - 115) * W:], text, W, 624, 115)
                    self.blend("SoftLight32AA", d, rt2, fparam=raker_text, n=N)
                elif raker_text >= 1.0:
                    raker_blur = _clampf(0.0, 100.0, T.getf("closeSpike:MoonrakerBlur"))
                    if raker_blur >= 1.0:
                        rt3 = self.buf(N)
                        self.old_blur("h", rt3, text, 624, 115, self.box_blur_scale(raker_blur))
                        text = rt3
                    self.blit("BlitSrc32", d[(H - 115) * W:], text, W, 624, 115)
                fade_flash(d, 0.0, fade_white)
                full("Overlay32", d, "assets/closeup/raker-LensDirt5_invert.png")
                fade_flash(d, fade_black, 0.0)
        elif dirt == 2:
            full("SoftLight32AA", d, "assets/greetings/Bokeh_Lens_Dirt_51.png", float(_f32(0.09) * K_GOLDEN_ANGLE))
        elif dirt == 3:
            full("SoftLight32AA", d, "assets/greetings/Bokeh_Lens_Dirt_51.png", float(_f32(0.075) * K_GOLDEN_ANGLE))
        if dirt != 1:
            fade_flash(d, fade_black, fade_white)
    elif part == 8:
        logo_idx = max(0, min(4, T.geti("demo:MainLogoIndex")))
        d = effect("spikey_distant", "spikey", close=False)
        fade_flash(d, fade_black, fade_white)
        full("SoftLight32", d, "assets/spikeball/SpikeyBall_byPass_BG_Overlay.png")
        full("Sub32", d, "assets/spikeball/Vignette_Layer02_inverted.png")
        full("Excl32", d, "assets/spikeball/nytrik-TheYearWas_Overlay_LensDirt.jpg")
        full("MulSrc32A", d, "assets/demo/tpb-06-dirty-vignette-1280x720.png")
        if logo_idx != 0:
            full("MixOver32", d, f"assets/spikeball/Layer 2023_{logo_idx}.png")
        full("Overlay32", d, "assets/spikeball/Vignette_CoolFilmLook.png")
    elif part == 9:
        d = effect("tunnel", "tunnel")
        full("Sub32", d, "assets/tunnels/Vignette_Layer02_inverted.png")
        show = _clampf(0.0, 3.0, T.getf("demo:Show2006"))
        if show > 0.0:
            self.blend("MixOver32", d, logo_blend(show, [f"assets/tunnels/layer 2006_{i}.png" for i in range(1, 5)], W, H), n=N)
    elif part == 10:
        overlay_a = _sat(T.getf("demo:WaterLove"))
        d = effect("sinuses", "sinuses")
        overlay = layer("assets/underwater/love prism_alpha 1280_720.png")
        blur = _clampf(0.0, 100.0, T.getf("demo:LoveBlurHorZ"))
        if blur != 0.0:
            rt0 = self.buf(N)
            self.old_blur("h", rt0, overlay, W, H, self.box_blur_scale(blur))
            overlay = rt0
        self.blit("BlitAdd32A", d, overlay, W, W, H, float(overlay_a))
        if T.geti("demo:Dirt") != 0:
            full("MulSrc32", d, "assets/underwater/LensDirt3_invert.png")
        fade_flash(d, fade_black, fade_white)
    elif part == 11:
        d = effect("laura", "laura")
        full("Darken32_50", d, f"assets/greetings/Greetings_Part{T.geti('demo:GreetSwitch') + 1}_BG_Overlay.png")
        full("SoftLight32", d, "assets/greetings/Bokeh_Lens_Dirt_51.png")
        self.blit("BlitSrc32", d[24 + (((H - 243) // 2) + 227) * W:], layer("assets/demo/tpb_xbox_tp-263x243.png"), W, 263, 243)
        full("Overlay32", d, "assets/greetings/Vignette_CoolFilmLook.png")
    elif part == 12:
        rt0 = self.buf(N); rt0[:] = 0xFFFFFF
        if T.geti("demo:FullWarpTPB") == 0:
            d = self.buf(N); d[:] = 0xFFFFFF
            rib_x = max(0, min(W, T.geti("demo:RibbonsX")))
            self.mix_src_s(d, layer("assets/demo/ribbons.png")[rib_x:], W, H - 1, 2160)
            kind = "h"
        else:
            d = effect("plasma", "plasma")
            kind = "v"
        full("MixSrc32", rt0, "assets/demo/TPB-logo.png")
        blur = T.getf("demo:BlurTPB")
        if blur != 0.0:
            self.old_blur(kind, rt0, rt0, W, H, self.box_blur_scale(blur))
        rt1 = self.buf(N)
        self.tape_warp(rt1, rt0, W, H, T.getf("demo:DistortStrengthTPB"), T.getf("demo:DistortTPB"))
        self.blend("MixOver32", d, rt1, n=N)
        full("MulSrc32", d, "assets/nautilus/Vignette.png")
    elif part == 13:
        d = self.buf(N)
        guys, joke = _sat(T.getf("demo:DiscoGuys")), _sat(T.getf("demo:CheapGPU"))
        if guys > 0.0:
            x_start, y_offs = (W - 8 * 128) >> 1, ((H - 128) >> 1) + 16
            for i, name in enumerate(("1", "1b", "2", "2b", "3", "3b", "4", "4b")):
                t = _sat(T.getf(f"demo:DiscoGuy{i + 1}"))
                s = _f32(_f32(_f32(t * t) * t) * _f32(_f32(t * _f32(_f32(t * _f32(6.0)) - _f32(15.0))) + _f32(10.0)))   # smootherstepf(0, 1, t)
                s = _f32(_f32(0.0) + _f32(_f32(_f32(1.0) - _f32(0.0)) * s))
                self.blit("BlitSrc32A", d[x_start + i * 128 + y_offs * W:], layer(f"assets/demo/tpb-06-disco-guy/{name}.png"), W, 128, 128, float(_f32(guys * s)))
                if guys < 1.0:
                    strip = d[y_offs * W:]
                    self.old_blur("h", strip, strip, W, 128, self.box_blur_scale(float(_f32(_f32(_f32(_f32(1.0) - guys) * K_2PI) * K_GOLDEN_ANGLE))))
            self.blit("BlitAdd32A", d[(((W - 1100) // 2) - 1) + (y_offs + 130) * W:], layer("assets/demo/are-we-done-1100x57.png"), W, 1100, 57, float(guys))
        elif joke > 0.0:
            self.blit("BlitSrc32A", d[((W - 960) // 2) + ((H - 160) // 2) * W:], layer("assets/demo/GPU-joke.png"), W, 960, 160, float(joke))
    else:
        raise ValueError(f"demo:Effect {part}: FxBlitter_DrawTestPattern is not restated")

    if part not in (1, 2, 3, 6, 7, 8, 10):
        fade_flash(d, fade_black, fade_white)
    return d[:N].reshape(H, W)


Port.demo_draw = _demo_draw
