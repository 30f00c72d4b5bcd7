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
