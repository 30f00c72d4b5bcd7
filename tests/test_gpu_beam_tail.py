"""The beam tail of the ball (ball.cpp:168-203) for EVERY tail length of a 3840- and a 1280-pixel row.

The reference accumulates `curStep += alphaStep` in float, pixel by pixel; the kernel walks that chain binade by binade
(csrc/ckd_voxel.cu, beam_tail).  The check is a numpy float32 restatement of the reference loop -- np.add.accumulate is the
same sequential rounded addition -- so every alphaStep = 1/(remainder-1) the two resolutions can produce is covered, ties of the
dropped bits included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def reference_tails(row_pixels, first, rows, beam_color, alpha_min, fill, raw_steps=False):
    out = np.full((rows, row_pixels), fill, dtype=np.uint32)
    col = np.uint32(beam_color & 0xFFFFFF)
    r, g, b = (beam_color >> 16) & 0xFF, (beam_color >> 8) & 0xFF, beam_color & 0xFF
    lum = np.float32(((r * 4731) >> 16) + ((g * 46871) >> 16) + ((b * 13932) >> 16))   # ball.cpp:181-188
    a_min = np.float32(alpha_min)
    three, two = np.float32(3), np.float32(2)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for row in range(rows):
            rem = first + row
            if rem == 0:
                continue
            step = np.float32(1) / np.float32(np.uint32(rem - 1))
            cur = np.zeros(rem, dtype=np.float32)
            if rem > 1:
                cur[1:] = np.add.accumulate(np.full(rem - 1, step, dtype=np.float32), dtype=np.float32)
            start = row_pixels - 1 - rem
            if raw_steps:
                out[row, start:start + rem] = cur.view(np.uint32)
                continue
            t = (cur * cur) * (three - two * cur)                    # smoothstepf, Math.h:59-63
            alpha = a_min + (lum - a_min) * t                        # lerpf, Math.h:52-56
            a = np.where(np.isfinite(alpha), alpha, 0).astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
            out[row, start:start + rem] = col | ((a << np.uint64(24)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return out


@pytest.mark.parametrize("row_pixels", [3840, 1280])
@pytest.mark.parametrize("beam_color,alpha_min", [(0x00C8B4A0, 0.0), (0xFFFFFFFF, 31.5), (0x00102030, 200.0), (0x00F0E010, -12.25)])
def test_every_tail_length(row_pixels, beam_color, alpha_min):
    from cookiedough_b200 import capi
    ctx = capi.Context(1280, 720, 0)
    try:
        rows, fill = row_pixels, 0x12345678
        d = ctx.malloc(rows * row_pixels * 4)
        ctx.upload(d, np.full((rows, row_pixels), fill, dtype=np.uint32))
        ctx.ball_beam_tail(d, row_pixels, rows, 0, beam_color, alpha_min)
        got = ctx.download(d, (rows, row_pixels))
        ctx.free(d)
    finally:
        ctx.close()
    want = reference_tails(row_pixels, 0, rows, beam_color, alpha_min, fill)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{len(bad)} pixels differ, first at row (= tail length) {bad[0][0]}, pixel {bad[0][1]}: {got[tuple(bad[0])]:#x} != {want[tuple(bad[0])]:#x}"


@pytest.mark.parametrize("row_pixels", [3840, 1280, 4099])
def test_accumulated_step_of_every_tail_length(row_pixels):
    """the chain itself, bit for bit: the pixels above only show it through a smoothstep and a truncation"""
    from cookiedough_b200 import capi
    ctx = capi.Context(1280, 720, 0)
    try:
        rows, fill = row_pixels, 0xFFFFFFFF
        d = ctx.malloc(rows * row_pixels * 4)
        ctx.upload(d, np.full((rows, row_pixels), fill, dtype=np.uint32))
        ctx.ball_beam_tail(d, row_pixels, rows, 0, 0, 0.0, raw_steps=True)
        got = ctx.download(d, (rows, row_pixels))
        ctx.free(d)
    finally:
        ctx.close()
    want = reference_tails(row_pixels, 0, rows, 0, 0.0, fill, raw_steps=True)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{len(bad)} values differ, first at tail length {bad[0][0]}, pixel {bad[0][1]}: {got[tuple(bad[0])]:#x} != {want[tuple(bad[0])]:#x}"


def test_tail_arguments_are_checked():
    from cookiedough_b200 import capi
    ctx = capi.Context(1280, 720, 0)
    try:
        d = ctx.malloc(64 * 4)
        with pytest.raises(capi.CkdError):
            ctx.ball_beam_tail(d, 64, 1, 64, 0, 0.0)     # a tail longer than its row
        ctx.free(d)
    finally:
        ctx.close()
