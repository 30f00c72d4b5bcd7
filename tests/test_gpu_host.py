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


def test_host_rect_blits_into_a_frame(host_and_ref):
    """BlitSrc32/A, BlitAdd32/A with pDest pointing INTO a frame (demo.cpp:886), MixSrc32S with a source wider than a frame
    (the 2160-pixel ribbons, demo.cpp:682), memset32, Polar_Blit_2x2 and FxBlitter_DrawTestPattern on host buffers"""
    host, R = host_and_ref
    import post_cases as pc
    from oracle.ref import aligned_u32
    n = R.res_x * R.res_y
    sprite_w, sprite_h = 263, 243
    sprite = aligned_u32(sprite_w * sprite_h, pad=64); sprite[:] = pc.seeded(sprite.size, "noise")
    offs = 101 + 57 * R.res_x

    def pair():
        a = aligned_u32(n, pad=4 * R.res_x); a[:] = pc.seeded(n, "mix")
        b = aligned_u32(n, pad=4 * R.res_x); b[:] = a
        return a, b

    for op, alpha in (("BlitSrc32", 0.0), ("BlitSrc32A", 0.6), ("BlitAdd32", 0.0), ("BlitAdd32A", 0.35)):
        a, b = pair()
        R.blit(op, a[offs:], sprite, R.res_x, sprite_w, sprite_h, alpha)
        host.post(op, b[offs:], sprite, R.res_x, sprite_w, f0=alpha, u=sprite_h)
        assert_bit_exact(b, a, op)

    ribbons = aligned_u32(2160 * (R.res_y - 1) + R.res_x, pad=64); ribbons[:] = pc.seeded(ribbons.size, "noise2")
    a, b = pair()
    R.mix_src_s(a, ribbons[77:], R.res_x, R.res_y - 1, 2160)
    host.post("MixSrc32S", b, ribbons[77:], R.res_x, R.res_y - 1, u=2160)
    assert_bit_exact(b, a, "MixSrc32S")

    a, b = pair()
    R.memset32(a, 0x00c0ffee, n - 8)
    host.post("memset32", b, None, a=n - 8, u=0x00c0ffee)
    assert_bit_exact(b, a, "memset32")

    nfx = R.fx_x * R.fx_y
    fx_src = aligned_u32(nfx, pad=4 * R.res_x); fx_src[:] = pc.seeded(nfx, "noise")
    for inverse in (False, True):
        a = aligned_u32(nfx, pad=4 * R.res_x); a[:] = 0x11223344
        b = aligned_u32(nfx, pad=4 * R.res_x); b[:] = 0x11223344
        R.polar_blit_2x2(a, fx_src, inverse)
        host.post("Polar_Blit_2x2", b, fx_src, u=int(inverse))
        assert_bit_exact(b, a, f"Polar_Blit_2x2 inverse={inverse}")

    a, b = pair()
    R.fx_test_pattern(a)
    host.post("FxBlitter_DrawTestPattern", b, None)
    assert_bit_exact(b, a, "FxBlitter_DrawTestPattern")


def test_host_module_setup_and_globals(host_and_ref):
    """Polar/BoxBlur/FxBlitter/Shared _Create/_Destroy and the globals they own: g_pFxMap, g_renderTarget as caller scratch
    for a demo.cpp-style chain on host buffers, g_gradientUnp16, Ball_GetBackground"""
    host, R = host_and_ref
    import post_cases as pc
    assert host.module("Polar") and host.module("BoxBlur") and host.module("FxBlitter")
    # Shared_Create loads the two TPB logos (shared-resources.cpp:27-34): the bare-effects host has not registered them
    logos = {"assets/demo/TPB-logo.png": (R.res_x, R.res_y), "assets/demo/tpb_xbox_tp-263x243.png": (263, 243)}
    for path, (w, h) in logos.items():
        host.register_image(path, pc.seeded(w * h, "noise2").reshape(h, w))
    assert host.module("Shared")
    try:
        n, nfx = R.res_x * R.res_y, R.fx_x * R.fx_y
        fx0 = host.global_array("g_pFxMap", 0, (nfx,))
        rt0 = host.global_array("g_renderTarget", 0, (n,))
        rt3 = host.global_array("g_renderTarget", 3, (n,))
        assert fx0 is not None and rt0 is not None and rt3 is not None
        assert np.array_equal(host.global_array("g_pXboxLogoTPB", 0, (263 * 243,)), pc.seeded(263 * 243, "noise2"))
        grad = host.global_array("g_gradientUnp16", 0, (256, 8), dtype=np.uint16)
        assert np.array_equal(grad[:, :4], np.repeat(np.arange(256, dtype=np.uint16)[:, None], 4, axis=1)) and not grad[:, 4:].any()
        assert np.array_equal(host.global_array("Ball_GetBackground", 0, (n,)), R.ball_background())

        # FX map -> frame -> inverse polar -> blurred in place, through the globals, against the reference on its own globals
        fx0[:] = pc.seeded(nfx, "noise")
        R.fxmap(0).ravel()[:nfx] = fx0
        host.post("Fx_Blit_2x2", rt0, fx0)
        host.post("Polar_Blit", rt3, rt0, u=1)
        host.post("BoxBlur32", rt3, rt3, R.res_x, R.res_y, 0.11)
        r0, r3 = R.render_target(0), R.render_target(3)
        R.fx_blit_2x2(r0, R.fxmap(0)); R.polar_blit(r3, r0, True); R.old_blur("hv", r3, r3, R.res_x, R.res_y, 0.11)
        assert_bit_exact(rt3, np.asarray(r3).ravel()[:n], "chain on g_renderTarget")
    finally:
        for name in ("Shared", "FxBlitter", "BoxBlur", "Polar"):
            host.module(name, create=False)
    assert host.global_array("g_renderTarget", 0, (4,)) is None and host.global_array("g_pFxMap", 0, (4,)) is None


def test_effect_created_from_image_files(host_and_ref, tmp_path):
    """X_Create with nothing registered: the host layer decodes the files itself (Image_Load32 / Image_Load8 without DevIL).
    The landscape's three maps are written out as PNG (lossless, so the reference sees the same pixels) under the names
    the reference loads, dropped from the registry, and Landscape_Create reads them back through host/ckd_image.cpp."""
    from PIL import Image
    host, R = host_and_ref
    paths = ["assets/scape/D17.png", "assets/scape/C17W-edit.png", "assets/scape/foggradient.jpg"]
    for path in paths:
        arr = R.assets[path]
        out = tmp_path / path
        out.parent.mkdir(parents=True, exist_ok=True)
        if arr.dtype == np.uint8:
            img = Image.fromarray(arr, "L")
        else:
            bgra = arr.view(np.uint8).reshape(arr.shape[0], arr.shape[1], 4)
            img = Image.fromarray(np.ascontiguousarray(bgra[..., [2, 1, 0, 3]]), "RGBA")
        with open(out, "wb") as f:       # a PNG stream under the reference's file name: the decoder goes by content
            img.save(f, "PNG")
    host.reload_from_files("landscape", paths, tmp_path)
    R.set_row(500); host.set_row(500)
    out = np.zeros((R.res_y, R.res_x), dtype=np.uint32)
    host.draw("landscape", out)
    assert_bit_exact(out, R.draw("landscape").copy(), "landscape from decoded files")
    with pytest.raises(Exception, match="Can not load image"):
        host.reload_from_files("landscape", paths[:1], tmp_path / "nowhere")
    host.reload_from_files("landscape", paths[:1], tmp_path)


@pytest.mark.parametrize("effect,row", [("plasma", 2600), ("sinuses", 7800), ("laura", 8900), ("spikey_distant", 3600), ("spikey_close", 6800),
                                        ("nautilus", 5700), ("tunnel", 4500), ("ball", 1500), ("ball", 2060), ("twister", 2008), ("tunnelscape", 4300), ("landscape", 500)])
def test_banded_readback_delivers_the_same_frame(host_and_ref, effect, row):
    """CkdHost_SetReadbackBands: a synchronous X_Draw into a page-locked buffer renders and copies in row bands (raymarchers
    without a post chain) or falls back to one copy (everything else): the frame in pDest is the same, bit for bit"""
    host, R = host_and_ref
    host.set_row(row)
    plain = np.zeros((R.res_y, R.res_x), dtype=np.uint32)
    host.set_readback_bands(0)
    host.draw(effect, plain)
    banded = np.zeros((R.res_y, R.res_x), dtype=np.uint32)
    host.pin(banded)
    try:
        for bands in (4, 3, 8, 2):
            banded[:] = 0x55aa55aa
            host.set_readback_bands(bands)
            host.draw(effect, banded)
            assert np.array_equal(banded, plain), f"{effect}@{row}: {bands} bands differ in {np.count_nonzero(banded != plain)} pixels"
        pageable = np.full((R.res_y, R.res_x), 0x55aa55aa, dtype=np.uint32)   # not page-locked: the arm is ignored, one copy
        host.draw(effect, pageable)
        assert np.array_equal(pageable, plain)
    finally:
        host.set_readback_bands(-1)
        host.unpin(banded)


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


def _fastcos_inputs():
    rng = np.random.default_rng(1234)
    edges = np.arange(0, 1025, dtype=np.float64) * (2.0 * np.pi / 1024.0)       # the table's phase boundaries ...
    near = np.concatenate([edges - 1e-9, edges, edges + 1e-9, np.nextafter(edges, np.inf), np.nextafter(edges, -np.inf)])
    return np.concatenate([
        near, -near, near + 2.0 * np.pi * 1000.0,
        rng.uniform(-1e6, 1e6, 200000), rng.uniform(-8.0, 8.0, 100000), rng.uniform(-1.0, 1.0, 50000),
        np.array([0.0, -0.0, 1.0, -1.0, 0.25, 1e-300, 5e-324, 1e9, -1e9, 1e12, 2.0**40, 2.0**52, 2.0**53 + 2, 1e15]),
    ])


@pytest.mark.parametrize("sine", [False, True], ids=["fastcosf", "fastsinf"])
def test_fastcosf_matches_reference(host_and_ref, sine):
    """fast-cosine.h:17-53 on the device (table in shared memory) against the reference's inline function, bit for bit:
    +-1e6, the PLL domain, every table edge and its neighbours, and arguments large enough that the exponent shift wraps"""
    host, R = host_and_ref
    x = _fastcos_inputs()
    got = host.fastcos(x, sine=sine)
    arg = x - 0.25 if sine else x
    want = np.array([R.lib.ref_fastcosf(float(v)) for v in arg], dtype=np.float32)
    bad = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
    assert bad.size == 0, f"{bad.size} of {x.size} differ, first x={x[bad[0]]!r}: {got[bad[0]]!r} vs {want[bad[0]]!r}"
    # the C ABI entry directly (device arrays) gives the same values
    assert np.array_equal(host.context().fastcos(x[:4096], sine=sine).view(np.uint32), want[:4096].view(np.uint32))


def test_fast_cos_table_is_the_reference_table(host_and_ref):
    host, R = host_and_ref
    host.fastcos(np.zeros(1))                        # InitializeFastCosine
    ref_tab = np.ctypeslib.as_array(R.lib.ref_fast_cos_tab(), shape=(1025,))
    assert np.array_equal(host.fast_cos_tab().view(np.uint64), ref_tab.view(np.uint64))
    assert np.array_equal(host.context().fast_cos_table().view(np.uint64), ref_tab.view(np.uint64))


def test_post_ops_in_place_on_one_host_buffer(host_and_ref):
    """ops that do not read their destination (TapeWarp32, Polar_Blit, Fx_Blit_2x2) called with pSrc == pDest must work
    on the caller's pixels (not on stale staging memory): the result equals the out-of-place call"""
    host, R = host_and_ref
    import post_cases as pc
    from oracle.ref import aligned_u32
    n = R.res_x * R.res_y
    src = pc.seeded(n, "noise")

    def buf(init):
        a = aligned_u32(n, pad=4 * R.res_x); a[:] = init
        return a

    for op, kw in (("TapeWarp32", dict(a=R.res_x, b=R.res_y, f0=0.5, f1=1.0)), ("Polar_Blit", dict(u=1)), ("Polar_Blit", dict(u=0))):
        want = buf(0)
        host.post(op, want, buf(src), **kw)
        host.post("MixSrc32", buf(pc.seeded(n, "mix")), buf(pc.seeded(n, "mix")), n)   # leaves other content in the staging targets
        got = buf(src)
        host.post(op, got, got, **kw)
        assert_bit_exact(got, want, f"{op} in place")
