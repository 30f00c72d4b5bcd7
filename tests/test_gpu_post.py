"""GPU parity of the integer 2D post chain: bit-exact against the committed goldens and the live compiled reference."""
import numpy as np
import pytest

import post_cases as pc
from util import assert_bit_exact, sha256_u32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", pc.CASES, ids=[c["label"] for c in pc.CASES])
def test_post_case_matches_golden(case, ctx_synth, golden_post):
    out = pc.run_cuda(ctx_synth, case)
    assert sha256_u32(out) == golden_post[case["label"]], case["label"]


def test_post_cases_match_live_reference(live720):
    R, ctx = live720
    for case in pc.CASES:
        assert_bit_exact(pc.run_cuda(ctx, case), pc.run_reference(R, case), case["label"])


def test_polar_maps_match_reference_behaviour(live720):
    """the UV maps are built on the host like the reference does: a remap of an index image pins every entry"""
    R, ctx = live720
    n = R.res_x * R.res_y
    src = np.arange(n, dtype=np.uint32).reshape(R.res_y, R.res_x) * np.uint32(0x01000193)
    for inverse in (False, True):
        ref_dst = R.frame()
        R.polar_blit(ref_dst, np.ascontiguousarray(src), inverse)
        d_src = ctx.to_device(src, pad_elems=4 * R.res_x)
        ctx.polar_blit(ctx.frame(), d_src, inverse)
        assert_bit_exact(ctx.read_frame(), ref_dst, f"polar inverse={inverse}")
        ctx.free(d_src)


# ---- 4K: size-independent properties + live reference -------------------------------------------------------------

def test_post_chain_4k_against_reference(live2160):
    """Polar_Blit(inverse) -> BoxBlur32 in place -> Fx_Blit_2x2 -> blends at 3840x2160"""
    R, ctx = live2160
    w, h = R.res_x, R.res_y
    n = w * h
    src = pc.seeded(n, "mul").reshape(h, w)
    dst = pc.seeded(n, "mix").reshape(h, w)
    fx = pc.seeded(R.fx_x * R.fx_y, "noise").reshape(R.fx_y, R.fx_x)

    from oracle.ref import aligned_u32
    r_src = aligned_u32(n, pad=4 * w).reshape(h, w); r_src[:] = src
    r_dst = aligned_u32(n, pad=4 * w).reshape(h, w); r_dst[:] = dst
    r_fx = aligned_u32(fx.size, pad=4 * w).reshape(fx.shape); r_fx[:] = fx
    r_tmp = aligned_u32(n, pad=4 * w).reshape(h, w)

    d_src = ctx.to_device(src, pad_elems=4 * w)
    d_dst = ctx.to_device(dst, pad_elems=4 * w)
    d_fx = ctx.to_device(fx, pad_elems=4 * w)
    d_tmp = ctx.to_device(np.zeros(n, dtype=np.uint32), pad_elems=4 * w)

    R.polar_blit(r_dst, r_src, True)
    ctx.polar_blit(d_dst, d_src, True)
    assert_bit_exact(ctx.download(d_dst, (h, w)), r_dst, "4K Polar_Blit(inverse)")

    for strength in (0.11, 0.01, 0.33, 1.0):
        R.old_blur("hv", r_dst, r_dst, w, h, strength)
        ctx.old_blur("hv", d_dst, d_dst, w, h, strength)
        assert_bit_exact(ctx.download(d_dst, (h, w)), r_dst, f"4K BoxBlur32 in place s={strength}")

    R.fx_blit_2x2(r_tmp, r_fx)
    ctx.fx_blit_2x2(d_tmp, d_fx)
    assert_bit_exact(ctx.download(d_tmp, (h, w)), r_tmp, "4K Fx_Blit_2x2")

    for op in ("MixSrc32", "SoftLight32", "Overlay32"):
        R.blend(op, r_dst, r_tmp)
        ctx.blend(op, d_dst, d_tmp, n)
        assert_bit_exact(ctx.download(d_dst, (h, w)), r_dst, f"4K {op}")

    R.new_blur("hv", r_tmp, r_dst, w, h, 6.28, 0.1, 3)
    ctx.new_blur("hv", d_tmp, d_dst, w, h, 6.28, 0.1, 3)
    assert_bit_exact(ctx.download(d_tmp, (h, w)), r_tmp, "4K BoxBlur_32 kGauss")

    R.polar_blit(r_dst, r_tmp, False, alpha=True)
    ctx.polar_blit(d_dst, d_tmp, False, alpha=True)
    assert_bit_exact(ctx.download(d_dst, (h, w)), r_dst, "4K Polar_BlitA")

    for d in (d_src, d_dst, d_fx, d_tmp):
        ctx.free(d)


def test_post_properties_4k(live2160):
    """identities that hold at any size (no oracle needed)"""
    _, ctx = live2160
    w, h = ctx.res_x, ctx.res_y
    n = w * h
    a = pc.seeded(n, "noise")
    const = np.full(n, 0x80402010, dtype=np.uint32)
    d_a = ctx.to_device(a, pad_elems=4 * w)
    d_b = ctx.to_device(a, pad_elems=4 * w)
    d_c = ctx.to_device(const, pad_elems=4 * w)
    d_z = ctx.to_device(np.zeros(n, dtype=np.uint32), pad_elems=4 * w)

    ctx.blend("Mix32", d_a, d_c, n, uparam=0)            # alpha 0 keeps dest
    assert np.array_equal(ctx.download(d_a, (n,)), a)
    ctx.blend("Add32", d_a, d_z, n)                       # + 0
    assert np.array_equal(ctx.download(d_a, (n,)), a)
    ctx.blend("Sub32", d_a, d_b, n)                       # x - x
    assert not ctx.download(d_a, (n,)).any()
    ctx.polar_blit(d_a, d_c, False)                       # remap of a constant image is constant
    assert np.array_equal(ctx.download(d_a, (n,)), const)
    fx = np.full(ctx.fx_x * ctx.fx_y, 0x11223344, dtype=np.uint32)
    d_fx = ctx.to_device(fx, pad_elems=4 * w)
    ctx.fx_blit_2x2(d_a, d_fx)                            # upsample of a constant map is constant
    assert np.array_equal(ctx.download(d_a, (n,)), np.full(n, 0x11223344, dtype=np.uint32))
    ctx.memset32(d_a, 0xdeadbeef, n)
    assert np.array_equal(ctx.download(d_a, (n,)), np.full(n, 0xdeadbeef, dtype=np.uint32))
    # blurring twice from the same input is deterministic
    ctx.upload(d_a, a); ctx.old_blur("hv", d_a, d_a, w, h, 0.2); first = ctx.download(d_a, (n,))
    ctx.upload(d_a, a); ctx.old_blur("hv", d_a, d_a, w, h, 0.2)
    assert np.array_equal(ctx.download(d_a, (n,)), first)
    for d in (d_a, d_b, d_c, d_z, d_fx):
        ctx.free(d)


def test_blend_chain_equals_sequential_blends(ctx_synth):
    """ckd_blend_chain: every op of the compositor's layer stacks in one pass == the same ops one launch at a time"""
    import numpy as np
    import post_cases as pc
    from cookiedough_b200 import capi
    ctx = ctx_synth
    n = 1280 * 720 - 3  # exercises the scalar tail as well
    layers = [ctx.to_device(pc.seeded(n + 3, name)) for name in ("noise", "noise2", "smooth")]
    start = pc.seeded(n + 3, "mix")
    ops = list(capi.BLEND_OPS)
    rng = np.random.default_rng(7)
    for trial in range(6):
        count = [1, 3, 8, 11, 15, 5][trial]
        steps = []
        for k in range(count):
            op = ops[(trial * 5 + k * 3) % len(ops)]
            src = None if op == "Fade32" else layers[int(rng.integers(0, 3))]
            f = 0.37 if op == "SoftLight32AA" else 0.0
            u = 77 if op == "Mix32" else ((200 << 24) | 0x123456) if op == "Fade32" else 0
            steps.append((op, src, f, u))
        d_seq = ctx.to_device(start)
        d_chain = ctx.to_device(start)
        if trial == 3:
            steps[4] = (steps[4][0] if steps[4][0] != "Fade32" else "Add32", d_chain, steps[4][2], steps[4][3])  # a source aliasing the destination
        for op, src, f, u in steps:
            ctx.blend(op, d_seq, d_seq if src == d_chain else (src or d_seq), n, f, u)
        ctx.blend_chain(d_chain, steps, n)
        a = ctx.download(d_seq, (n + 3,))
        b = ctx.download(d_chain, (n + 3,))
        assert np.array_equal(a, b), f"trial {trial}: {[s[0] for s in steps]}"
        assert np.array_equal(b[n:], start[n:]), "wrote past num_pixels"


@pytest.mark.parametrize("kind", ["h", "v"])
def test_old_blur_every_kernel_size_in_place(live720, kind):
    """every kernel width 1..255 (each one picks its own kernel: register trailing edge, serial walk, or a scan shape P x QT)
    against the live reference, in place, on images whose lines are just long enough for the widest kernel"""
    from oracle.ref import aligned_u32
    R, ctx = live720
    w, h = (704, 40) if kind == "h" else (64, 520)
    n = w * h
    src = pc.seeded(n, "noise") | np.uint32(0x40404040)   # bright enough to reach the 16-bit saturation for wide kernels
    d = ctx.to_device(src, pad_elems=4 * w)
    bad = []
    for span in range(1, 256):
        strength = span / 255.0
        ref = aligned_u32(n, pad=4 * w)
        ref[:] = src
        R.old_blur(kind, ref, ref, w, h, strength)
        ctx.upload(d, src)
        ctx.old_blur(kind, d, d, w, h, strength)
        if not np.array_equal(ctx.download(d, (n,)), ref):
            bad.append(span)
    ctx.free(d)
    assert not bad, f"old_blur_{kind} in place differs for kernel spans {bad}"
