"""The drop-in boundary end to end: the C++ host layer's reference-named entry points (X_Draw into a HOST buffer, Rocket
read inside the call) against the compiled reference driven at the same pinned time."""
import numpy as np
import pytest

from util import INTEGER_EFFECTS, assert_bit_exact, assert_float_parity, seed_frame

pytestmark = pytest.mark.gpu

CASES = [("plasma", 2600), ("nautilus", 5700), ("spikey_close", 6800), ("spikey_distant", 3600), ("tunnel", 4500), ("sinuses", 7800),
         ("laura", 8900), ("landscape", 500), ("tunnelscape", 4300), ("ball", 1500), ("ball", 2060), ("twister", 2008), ("landscape", 1030)]


@pytest.fixture(scope="module")
def host_and_ref():
    from oracle import ref as oref
    if not oref.available(720):
        pytest.skip("oracle/_ref not built")
    from cookiedough_b200 import hostapi
    from cookiedough_b200.assets import Assets
    R = oref.Reference.get(720, Assets(1280, 720))
    host = hostapi.Host(1280, 720, 0, R.assets)
    yield host, R
    host.close()


@pytest.mark.parametrize("effect,row", CASES, ids=[f"{e}@{r}" for e, r in CASES])
def test_host_draw_matches_reference(host_and_ref, effect, row):
    host, R = host_and_ref
    R.set_row(row)
    host.set_row(row)
    seed = seed_frame(R.res_x, R.res_y)
    R.render_target(0)[:] = seed
    ctx = host.context()
    ctx.upload(ctx.render_target(0), seed)
    ref_out = R.draw(effect).copy()
    out = np.zeros((R.res_y, R.res_x), dtype=np.uint32)
    host.draw(effect, out)
    kind = "ball" if effect == "ball" else effect
    if kind in INTEGER_EFFECTS:
        assert_bit_exact(out, ref_out, f"{effect}@{row}")
    else:
        assert_float_parity(out, ref_out, f"{effect}@{row}")


def test_host_post_ops_on_host_buffers(host_and_ref):
    host, R = host_and_ref
    import post_cases as pc
    from oracle.ref import aligned_u32
    n = R.res_x * R.res_y
    src = pc.seeded(n, "noise")
    dst = pc.seeded(n, "mix")

    def pair():
        a = aligned_u32(n, pad=4 * R.res_x); a[:] = dst
        b = aligned_u32(n, pad=4 * R.res_x); b[:] = dst
        return a, b

    s = aligned_u32(n, pad=4 * R.res_x); s[:] = src
    a, b = pair(); R.polar_blit(a, s, True); host.post("Polar_Blit", b, s, u=1); assert_bit_exact(b, a, "Polar_Blit")
    a, b = pair(); R.polar_blit(a, s, False, alpha=True); host.post("Polar_BlitA", b, s, u=0); assert_bit_exact(b, a, "Polar_BlitA")
    a, b = pair(); R.old_blur("hv", a, a, R.res_x, R.res_y, 0.11); host.post("BoxBlur32", b, b, R.res_x, R.res_y, 0.11); assert_bit_exact(b, a, "BoxBlur32 in place")
    a, b = pair(); R.old_blur("h", a, s, R.res_x, R.res_y, 0.3); host.post("HorizontalBoxBlur32", b, s, R.res_x, R.res_y, 0.3); assert_bit_exact(b, a, "HorizontalBoxBlur32")
    a, b = pair(); R.blend("MixSrc32", a, s); host.post("MixSrc32", b, s, n); assert_bit_exact(b, a, "MixSrc32")
    a, b = pair(); R.tape_warp(a, s, R.res_x, R.res_y, 0.5, 0.33); host.post("TapeWarp32", b, s, R.res_x, R.res_y, 0.5, 0.33); assert_bit_exact(b, a, "TapeWarp32")
    fx = aligned_u32(R.fx_x * R.fx_y, pad=4 * R.res_x); fx[:] = pc.seeded(fx.size, "noise2")
    a, b = pair(); R.fx_blit_2x2(a, fx); host.post("Fx_Blit_2x2", b, fx); assert_bit_exact(b, a, "Fx_Blit_2x2")


def test_pinning_the_callers_frame_buffer(host_and_ref):
    """CkdHost_PinFrameBuffer: same frame, faster copy-back into a caller-owned (malloc'ed) buffer"""
    import time
    host, R = host_and_ref
    host.set_row(2600)
    plain = np.zeros((R.res_y, R.res_x), dtype=np.uint32)
    pinned = np.zeros((R.res_y, R.res_x), dtype=np.uint32)
    host.draw("plasma", plain)
    host.pin(pinned)
    try:
        host.draw("plasma", pinned)
        assert np.array_equal(plain, pinned)
        def rate(buf):
            t0 = time.perf_counter()
            for _ in range(20):
                host.draw("plasma", buf)
            return time.perf_counter() - t0
        rate(plain); rate(pinned)
        assert rate(pinned) <= rate(plain) * 1.25  # never slower (typically 1.5-3x faster at 4K; 720p frames are small)
    finally:
        host.unpin(pinned)
