"""The arithmetic behind beam_tail (csrc/ckd_voxel.cu), restated in numpy float32 and checked on the CPU against the
reference's loop `curStep += alphaStep` (ball.cpp:195-203) for every tail length a 4K row can have and a few far longer ones.

While the running value stays inside one binade its grid is fixed, so a rounded addition of the constant adds a constant
multiple of that grid -- after the first addition made inside the binade (a tie rounds to even and can make that one differ).
The walk below is the kernel's, statement by statement; tests/test_gpu_beam_tail.py checks the kernel itself on the device."""
import numpy as np
import pytest

f32 = np.float32


def sequential(remainder):
    step = f32(1) / f32(np.uint32(remainder - 1))
    cur = np.zeros(remainder, dtype=np.float32)
    if remainder > 1:
        cur[1:] = np.add.accumulate(np.full(remainder - 1, step, dtype=np.float32), dtype=np.float32)
    return cur


def binade_walk(remainder):
    a = f32(1) / f32(np.uint32(remainder - 1))
    out = np.zeros(remainder, dtype=np.float32)
    # pixels 0..31: literal additions
    c = f32(0)
    for i in range(32):
        if i < remainder:
            out[i] = c
        c = f32(c + a)
    k = 32
    binades = 0
    while k < remainder:
        c1 = f32(c + a)
        c2 = f32(c1 + a)
        e = int(c.view(np.uint32)) >> 23
        base, inc, n = c, f32(0), 1
        if (int(c1.view(np.uint32)) >> 23) == e:
            inc = f32(c2 - c1)
            base = f32(c1 - inc)
            per_ulp = np.uint32((277 - e) << 23).view(np.float32)          # 2^(150 - e)
            base_units, inc_units = int(f32(base * per_ulp)), int(f32(inc * per_ulp))
            assert f32(base_units) == f32(base * per_ulp) and inc_units >= 1   # exact integers of the binade's grid
            n = 1 + (0xFFFFFF - base_units) // inc_units
        n = min(n, remainder - k)
        m = np.arange(n, dtype=np.uint32)
        vals = (base + m.astype(np.float32) * inc).astype(np.float32)        # one rounded multiply, one rounded add
        vals[0] = c
        out[k:k + n] = vals
        last = n - 1
        c = f32((c if last == 0 else f32(base + f32(last) * inc)) + a)
        k += n
        binades += 1
    return out, binades


@pytest.mark.parametrize("lo,hi", [(1, 1400), (1400, 2800), (2800, 4200)])
def test_walk_equals_sequential_accumulation(lo, hi):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for remainder in range(lo, hi):
            got, _ = binade_walk(remainder)
            want = sequential(remainder)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"tail length {remainder}: first difference at pixel {int(np.argmax(got.view(np.uint32) != want.view(np.uint32)))}"


@pytest.mark.parametrize("remainder", [16385, 65537, 100003, 262145, 1000003])
def test_walk_on_long_rows(remainder):
    got, binades = binade_walk(remainder)
    want = sequential(remainder)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert binades <= 24          # a binade per doubling of the running value, whatever the length


def test_ties_are_covered():
    """steps whose dropped bits are exactly half a grid unit in some binade the walk passes through (ties-to-even): present
    among the 4K tail lengths, so the sweep above exercises the first-step rule"""
    ties = 0
    for remainder in range(34, 4200):
        a = f32(1) / f32(remainder - 1)
        mant = int(a.view(np.uint32)) & 0x7FFFFF | 0x800000
        ea = (int(a.view(np.uint32)) >> 23)
        for e in range(ea + 1, 127):                      # binades above a's own, up to [0.5, 1)
            drop = e - ea                                  # bits of a below the binade's grid
            if drop <= 24 and (mant & ((1 << drop) - 1)) == (1 << (drop - 1)):
                ties += 1
                break
    assert ties > 0
