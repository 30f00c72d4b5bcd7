import hashlib

import numpy as np

# north star tolerance for the float (raymarch) paths: <= 2 LSB per channel and >= 99.5 % of the pixels exact
MAX_LSB = 2
MIN_EXACT_PCT = 99.5

INTEGER_EFFECTS = {"landscape", "tunnelscape", "ball", "twister"}


def sha256_u32(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).astype("<u4").tobytes()).hexdigest()


def pixel_stats(a, b):
    a8 = np.ascontiguousarray(a).view(np.uint8).reshape(-1, 4).astype(np.int16)
    b8 = np.ascontiguousarray(b).view(np.uint8).reshape(-1, 4).astype(np.int16)
    d = np.abs(a8 - b8).max(axis=1)
    return 100.0 * float((d == 0).sum()) / d.size, int(d.max())


def assert_float_parity(out, ref, what):
    exact, max_delta = pixel_stats(out, ref)
    assert max_delta <= MAX_LSB and exact >= MIN_EXACT_PCT, f"{what}: {exact:.4f}% exact, max delta {max_delta} LSB"


def assert_bit_exact(out, ref, what):
    if not np.array_equal(out, ref):
        exact, max_delta = pixel_stats(out, ref)
        raise AssertionError(f"{what}: not bit-exact ({exact:.4f}% exact, max delta {max_delta})")


def seed_frame(res_x, res_y):
    n = res_x * res_y
    return (np.arange(n, dtype=np.uint32) * np.uint32(2654435761)).reshape(res_y, res_x)
