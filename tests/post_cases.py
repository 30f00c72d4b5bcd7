"""Shared definitions of the 2D post-chain parity cases (SURVEY.md section 8d config 4 + the edge cases the demo hits).

Used by tests/golden/make_golden.py (reference side, to pin hashes), tests/test_ref_golden.py (reference vs pins),
tests/test_gpu_post.py (CUDA vs pins and vs the live reference) -- one list, three consumers.
Inputs are generated with integer hashes only, so they are identical on every machine.
"""
import numpy as np

RES_X, RES_Y = 1280, 720
FX_X, FX_Y = RES_X // 2 + 4, RES_Y // 2 + 4


def _hash_u32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def seeded(n, kind):
    i = np.arange(n, dtype=np.uint32)
    if kind == "mul":      # src[i] = i*2654435761
        return i * np.uint32(2654435761)
    if kind == "mix":      # dst[i] = i*40503 + (i<<20)
        return i * np.uint32(40503) + (i << np.uint32(20))
    if kind == "noise":
        return _hash_u32(i + np.uint32(1234))
    if kind == "noise2":
        return _hash_u32(i * np.uint32(3) + np.uint32(99991))
    if kind == "smooth":   # slowly varying image: blur accumulators stay far from saturation
        x = (i % np.uint32(RES_X)).astype(np.uint32)
        y = (i // np.uint32(RES_X)).astype(np.uint32)
        b = (x >> np.uint32(2)) & np.uint32(0xff)
        g = (y >> np.uint32(1)) & np.uint32(0xff)
        r = ((x + y) >> np.uint32(3)) & np.uint32(0xff)
        a = np.uint32(255) - b
        return b | (g << np.uint32(8)) | (r << np.uint32(16)) | (a << np.uint32(24))
    if kind == "bright":   # saturates the 16-bit blur accumulators for wide kernels when blurred in place
        return _hash_u32(i + np.uint32(7)) | np.uint32(0xc0c0c0c0)
    raise ValueError(kind)


CASES = []


def _add(**kw):
    CASES.append(kw)


_add(label="fx_blit_2x2/noise", op="fx_blit", src="noise")
_add(label="fx_blit_2x2/mul", op="fx_blit", src="mul")
for inv in (0, 1):
    for alpha in (0, 1):
        _add(label=f"polar_blit/inv{inv}/alpha{alpha}", op="polar", inverse=inv, alpha=alpha, src="noise", dst="mix")
for inv in (0, 1):
    _add(label=f"polar_blit_2x2/inv{inv}", op="polar_2x2", inverse=inv, src="noise", dst="mix")
_add(label="fx_test_pattern", op="test_pattern", dst="mix")

for kind in ("h", "v", "hv"):
    for strength in (0.01, 0.05, 0.11, 0.33, 1.0):
        for inplace in (True, False):
            _add(label=f"old_blur_{kind}/s{strength}/{'inplace' if inplace else 'copy'}", op="old_blur", kind=kind, w=RES_X, h=RES_Y,
                 strength=strength, inplace=inplace, src="noise", dst="mix")
_add(label="old_blur_hv/bright/s1.0/inplace", op="old_blur", kind="hv", w=RES_X, h=RES_Y, strength=1.0, inplace=True, src="bright", dst="mix")
_add(label="old_blur_hv/smooth/s0.33/inplace", op="old_blur", kind="hv", w=RES_X, h=RES_Y, strength=0.33, inplace=True, src="smooth", dst="mix")
# the other buffer shapes the demo blurs: FX map, 1280x568 credits, 624x115, a 1280x128 strip
for (w, h) in ((FX_X, FX_Y), (1280, 568), (624, 115), (1280, 128)):
    for kind in ("h", "v", "hv"):
        _add(label=f"old_blur_{kind}/{w}x{h}/s0.2", op="old_blur", kind=kind, w=w, h=h, strength=0.2, inplace=True, src="noise", dst="mix")
_add(label="old_blur_h/span1", op="old_blur", kind="h", w=RES_X, h=64, strength=0.004, inplace=True, src="noise", dst="mix")
_add(label="old_blur_v/span255", op="old_blur", kind="v", w=256, h=300, strength=2.0, inplace=False, src="noise", dst="mix")

for kind in ("h", "v", "hv"):
    for passes in (1, 2, 3):
        _add(label=f"new_blur_{kind}/p{passes}/s6.28", op="new_blur", kind=kind, w=RES_X, h=RES_Y, strength=6.28, gain=0.1, passes=passes, src="noise", dst="mix")
_add(label="new_blur_hv/p3/s30/g0", op="new_blur", kind="hv", w=RES_X, h=RES_Y, strength=30.0, gain=0.0, passes=3, src="smooth", dst="mix")
_add(label="new_blur_h/p1/s100/g1", op="new_blur", kind="h", w=RES_X, h=RES_Y, strength=100.0, gain=1.0, passes=1, src="noise", dst="mix")
_add(label="new_blur_hv/fxmap/p2/s12.5", op="new_blur", kind="hv", w=FX_X, h=FX_Y, strength=12.5, gain=0.25, passes=2, src="noise2", dst="mix")
_add(label="new_blur_h/p1/s0.7", op="new_blur", kind="h", w=RES_X, h=RES_Y, strength=0.7, gain=0.0, passes=1, src="noise", dst="mix")

_add(label="tape_warp/0.5/0.33", op="tape_warp", strength=0.5, speed=0.33, src="noise", dst="mix")
_add(label="tape_warp/landscape", op="tape_warp", strength=0.02, speed=0.33, src="smooth", dst="mix")
_add(label="tape_warp/strong", op="tape_warp", strength=40.0, speed=0.013, src="noise2", dst="mix")

_BLEND_ARGS = {"Mix32": (0.0, 77), "SoftLight32AA": (0.37, 0), "Fade32": (0.0, (200 << 24) | 0x123456)}
for op in ("Mix32", "MixOver32", "Add32", "Sub32", "Excl32", "SoftLight32", "SoftLight32A", "SoftLight32AA", "Overlay32", "Overlay32A",
           "Darken32_50", "MulSrc32", "MulSrc32A", "MixSrc32", "Fade32"):
    f, u = _BLEND_ARGS.get(op, (0.0, 0))
    _add(label=f"blend/{op}/full", op="blend", blend=op, n=RES_X * RES_Y, fparam=f, uparam=u, src="noise", dst="noise2")
    _add(label=f"blend/{op}/ragged", op="blend", blend=op, n=12345, fparam=f, uparam=u, src="mul", dst="mix")
_add(label="blend/Mix32/alpha0", op="blend", blend="Mix32", n=4096, fparam=0.0, uparam=0, src="noise", dst="noise2")
_add(label="blend/Mix32/alpha255", op="blend", blend="Mix32", n=4096, fparam=0.0, uparam=255, src="noise", dst="noise2")
_add(label="blend/SoftLight32AA/alpha1.5", op="blend", blend="SoftLight32AA", n=4096, fparam=1.5, uparam=0, src="noise", dst="noise2")
_add(label="blend/Add32/empty", op="blend", blend="Add32", n=0, fparam=0.0, uparam=0, src="noise", dst="noise2")

for op in ("BlitSrc32", "BlitSrc32A", "BlitAdd32", "BlitAdd32A"):
    for alpha in (0.6, 1.0, 0.0):
        _add(label=f"blit/{op}/sprite/a{alpha}", op="blit", blit=op, dest_res_x=RES_X, src_res_x=442, y_res=152, alpha=alpha, src="noise", dst="noise2")
    _add(label=f"blit/{op}/full", op="blit", blit=op, dest_res_x=RES_X, src_res_x=RES_X, y_res=RES_Y, alpha=0.85, src="mul", dst="mix")
_add(label="mix_src_s/ribbons", op="mix_src_s", dest_res_x=RES_X, dest_res_y=300, src_stride=2160, src="noise", dst="noise2")
_add(label="memset32", op="memset32", value=0x80c0ffee, n=RES_X * 100, dst="mix")


def test_pattern():
    """what FxBlitter_DrawTestPattern writes into g_pFxMap[0] (fx-blitter.cpp:77-95): line stripes above, column stripes below"""
    y, x = np.mgrid[0:FX_Y, 0:FX_X]
    on = np.where(y < FX_Y // 2, y & 1, x & 1).astype(bool)
    return np.where(on, np.uint32(0xffffffff), np.uint32(0)).astype(np.uint32).ravel()


def _sizes(case):
    """(dst elements, src elements) each case needs"""
    op = case["op"]
    full = RES_X * RES_Y
    if op == "fx_blit":
        return full, FX_X * FX_Y
    if op in ("polar", "tape_warp"):
        return full, full
    if op == "polar_2x2":
        return FX_X * FX_Y, FX_X * FX_Y
    if op == "test_pattern":
        return full, 1
    if op in ("old_blur", "new_blur"):
        return case["w"] * case["h"], case["w"] * case["h"]
    if op == "blend":
        return max(case["n"], 1), max(case["n"], 1)
    if op == "blit":
        return case["dest_res_x"] * case["y_res"], case["src_res_x"] * case["y_res"]
    if op == "mix_src_s":
        return case["dest_res_x"] * case["dest_res_y"], case["src_stride"] * case["dest_res_y"]
    if op == "memset32":
        return case["n"], 1
    raise ValueError(op)


def inputs(case):
    nd, ns = _sizes(case)
    dst = seeded(nd, case.get("dst", "mix"))
    src = seeded(ns, case.get("src", "noise"))
    return dst, src


def run_reference(R, case):
    """runs the case on the compiled reference (oracle.ref.Reference) and returns the destination buffer"""
    from oracle.ref import aligned_u32
    dst0, src0 = inputs(case)
    pad = 4 * RES_X
    dst = aligned_u32(dst0.size, pad=pad)
    dst[:] = dst0
    src = aligned_u32(src0.size, pad=pad)
    src[:] = src0
    op = case["op"]
    if op == "fx_blit":
        R.fx_blit_2x2(dst, src)
    elif op == "polar":
        R.polar_blit(dst, src, bool(case["inverse"]), alpha=bool(case["alpha"]))
    elif op == "polar_2x2":
        R.polar_blit_2x2(dst, src, bool(case["inverse"]))
    elif op == "test_pattern":
        R.fx_test_pattern(dst)
    elif op == "old_blur":
        s = dst if case["inplace"] else src
        if case["inplace"]:
            dst[:] = src0
        R.old_blur(case["kind"], dst, s, case["w"], case["h"], case["strength"])
    elif op == "new_blur":
        # HorzBlur32 reads 2 pixels past the end of every line (boxblur.cpp:171,185).  Past the last line of its internal
        # ping-pong buffers that is whatever an earlier call left there: scrub them with a blur of a zero image so the
        # "virtual pixels after the end" are 0, which is also what the CUDA path defines (include/ckd.h).
        zeros = aligned_u32(RES_X * (RES_Y + 8), pad=pad)
        R.new_blur("hv", zeros, zeros, RES_X, RES_Y + 8, 1.0, 0.0, 2)
        R.new_blur("hv", zeros, zeros, RES_X, RES_Y + 8, 1.0, 0.0, 3)
        R.new_blur(case["kind"], dst, src, case["w"], case["h"], case["strength"], case["gain"], case["passes"])
    elif op == "tape_warp":
        R.tape_warp(dst, src, RES_X, RES_Y, case["strength"], case["speed"])
    elif op == "blend":
        if case["n"] > 0:
            R.blend(case["blend"], dst, src, case["fparam"], case["uparam"], n=case["n"])
    elif op == "blit":
        R.blit(case["blit"], dst, src, case["dest_res_x"], case["src_res_x"], case["y_res"], case["alpha"])
    elif op == "mix_src_s":
        R.mix_src_s(dst, src, case["dest_res_x"], case["dest_res_y"], case["src_stride"])
    elif op == "memset32":
        R.memset32(dst, case["value"], case["n"])
    else:
        raise ValueError(op)
    return dst.copy()


def run_cuda(ctx, case):
    """runs the case through the C ABI (cookiedough_b200.capi.Context) and returns the destination buffer"""
    import ctypes as C
    dst0, src0 = inputs(case)
    pad = 4 * RES_X
    op = case["op"]
    if op == "old_blur" and case["inplace"]:
        dst0 = src0.copy()
    d_dst = ctx.to_device(dst0, pad_elems=pad)
    d_src = ctx.to_device(src0, pad_elems=pad)
    try:
        if op == "fx_blit":
            ctx.fx_blit_2x2(d_dst, d_src)
        elif op == "polar":
            ctx.polar_blit(d_dst, d_src, bool(case["inverse"]), alpha=bool(case["alpha"]))
        elif op == "polar_2x2":
            ctx.polar_blit_2x2(d_dst, d_src, bool(case["inverse"]))
        elif op == "test_pattern":
            ctx.upload(ctx.fxmap(0), test_pattern())
            ctx.fx_blit_2x2(d_dst, ctx.fxmap(0))
        elif op == "old_blur":
            ctx.old_blur(case["kind"], d_dst, d_dst if case["inplace"] else d_src, case["w"], case["h"], case["strength"])
        elif op == "new_blur":
            ctx.new_blur(case["kind"], d_dst, d_src, case["w"], case["h"], case["strength"], case["gain"], case["passes"])
        elif op == "tape_warp":
            ctx.tape_warp(d_dst, d_src, RES_X, RES_Y, case["strength"], case["speed"])
        elif op == "blend":
            ctx.blend(case["blend"], d_dst, d_src, case["n"], case["fparam"], case["uparam"])
        elif op == "blit":
            ctx.blit(case["blit"], d_dst, d_src, case["dest_res_x"], case["src_res_x"], case["y_res"], case["alpha"])
        elif op == "mix_src_s":
            ctx._check(ctx.L.ckd_mix_src_s(ctx.h, C.c_void_p(d_dst), C.c_void_p(d_src), case["dest_res_x"], case["dest_res_y"], case["src_stride"]))
        elif op == "memset32":
            ctx.memset32(d_dst, case["value"], case["n"])
        else:
            raise ValueError(op)
        return ctx.download(d_dst, (dst0.size,))
    finally:
        ctx.free(d_dst)
        ctx.free(d_src)
