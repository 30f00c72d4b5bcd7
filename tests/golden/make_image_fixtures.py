#!/usr/bin/env python3
"""tests/golden/make_image_fixtures.py -- small PNG / JPEG files for the host layer's decoders (host/ckd_image.cpp).

The files are written with Pillow from integer-hash pixels; the expected pixels are what Pillow (libpng / libjpeg-turbo,
the decoders the reference's DevIL links as well) reads back, stored next to them in expected.npz as BGRA uint32 and L8.
Covers every PNG colour type and bit depth, tRNS in its three forms, Adam7; JPEG baseline and progressive at 4:4:4,
4:2:2 and 4:2:0, grey, restart intervals, odd sizes (partial MCUs) and a 2-pixel-wide chroma plane."""
import os

import numpy as np
from PIL import Image

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "images")


def _hash(n, seed):
    x = (np.arange(n, dtype=np.uint32) + np.uint32(seed)) * np.uint32(2654435761)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(13)
    return x


def noise(h, w, c, seed):
    return (_hash(h * w * c, seed) >> np.uint32(11)).astype(np.uint8).reshape(h, w, c)


def smooth(h, w, seed):
    """photo-like content (JPEG on noise is all clamping): gradients + a little texture"""
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    t = noise(h, w, 3, seed).astype(np.int64) >> 4
    r = (x * 5 + y * 2 + t[..., 0]) % 256
    g = (x * 2 + y * 7 + t[..., 1] + 64) % 256
    b = ((x + y) * 3 + t[..., 2] + 128) % 256
    return np.stack([r, g, b], axis=-1).astype(np.uint8)


def expected(path):
    img = Image.open(path)
    rgba = np.asarray(img.convert("RGBA"), dtype=np.uint8)
    bgra = np.ascontiguousarray(rgba[..., [2, 1, 0, 3]]).view(np.uint32).reshape(rgba.shape[0], rgba.shape[1])
    gray = np.asarray(Image.open(path).convert("L"), dtype=np.uint8)
    return bgra, gray


def main():
    os.makedirs(HERE, exist_ok=True)
    files = {}

    def png(name, img, **kw):
        img.save(os.path.join(HERE, name), "PNG", **kw)
        files[name] = None

    def jpg(name, arr, mode="RGB", **kw):
        Image.fromarray(arr, mode).save(os.path.join(HERE, name), "JPEG", **kw)
        files[name] = None

    h, w = 37, 53
    png("rgb8.png", Image.fromarray(noise(h, w, 3, 1), "RGB"))
    png("rgba8.png", Image.fromarray(noise(h, w, 4, 2), "RGBA"))
    png("grey8.png", Image.fromarray(noise(h, w, 1, 4)[..., 0], "L"))
    png("greyalpha8.png", Image.fromarray(noise(h, w, 2, 5), "LA"))
    png("grey1.png", Image.fromarray((noise(h, w, 1, 6)[..., 0] & 1) * 255, "L").convert("1"))
    pal = Image.fromarray(noise(h, w, 1, 7)[..., 0] % 200, "P")
    pal.putpalette(list(noise(1, 256, 3, 8).ravel()))
    png("palette8.png", pal)
    png("palette8_trns.png", pal, transparency=bytes((i * 7) % 256 for i in range(120)))
    png("palette8_trns_index.png", pal, transparency=0)
    pal4 = Image.fromarray(noise(h, w, 1, 9)[..., 0] % 16, "P")
    pal4.putpalette(list(noise(1, 16, 3, 10).ravel()))
    png("palette4.png", pal4, bits=4)
    pal2 = Image.fromarray(noise(h, w, 1, 11)[..., 0] % 4, "P")
    pal2.putpalette(list(noise(1, 4, 3, 12).ravel()))
    png("palette2.png", pal2, bits=2)
    key = noise(h, w, 3, 13)
    key[5:9, 7:20] = (12, 200, 77)
    png("rgb8_colourkey.png", Image.fromarray(key, "RGB"), transparency=(12, 200, 77))
    g16 = (_hash(h * w, 14) >> np.uint32(9)).astype(np.uint16).reshape(h, w)
    png("grey16.png", Image.fromarray(g16.astype(np.int32), "I").convert("I;16"))
    png("rgb8_filters.png", Image.fromarray(smooth(96, 80, 15), "RGB"), compress_level=9)   # smooth content: the encoder picks Sub/Up/Average/Paeth

    # Adam7 + 16-bit RGB(A): Pillow cannot write them; the raw PNG writer below can
    import struct
    import zlib

    def raw_png(name, samples, depth, color_type, interlace):
        hh, ww, cc = samples.shape
        def pack(img):
            rows = []
            for row in img:
                data = row.astype(">u2").tobytes() if depth == 16 else row.astype(np.uint8).tobytes()
                rows.append(b"\x00" + data)                     # filter 0; the filters are exercised by the Pillow-written files
            return b"".join(rows)
        if interlace:
            x0, y0, dx, dy = (0, 4, 0, 2, 0, 1, 0), (0, 0, 4, 0, 2, 0, 1), (8, 8, 4, 4, 2, 2, 1), (8, 8, 8, 4, 4, 2, 2)
            body = b"".join(pack(samples[y0[i]::dy[i], x0[i]::dx[i]]) for i in range(7) if samples[y0[i]::dy[i], x0[i]::dx[i]].size)
        else:
            body = pack(samples)
        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
        data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", ww, hh, depth, color_type, 0, 0, 1 if interlace else 0))
        data += chunk(b"IDAT", zlib.compress(body, 6)[:1000]) + chunk(b"IDAT", zlib.compress(body, 6)[1000:]) + chunk(b"IEND", b"")
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(data)
        files[name] = None

    raw_png("rgba8_adam7.png", noise(h, w, 4, 3), 8, 6, True)
    raw_png("rgb8_adam7_tiny.png", noise(3, 5, 3, 16), 8, 2, True)
    raw_png("rgb16.png", (_hash(h * w * 3, 17) >> np.uint32(7)).astype(np.uint16).reshape(h, w, 3), 16, 2, False)
    raw_png("rgba16_adam7.png", (_hash(h * w * 4, 18) >> np.uint32(7)).astype(np.uint16).reshape(h, w, 4), 16, 6, True)

    photo = smooth(67, 91, 20)
    jpg("base_444.jpg", photo, quality=90, subsampling=0)
    jpg("base_422.jpg", photo, quality=85, subsampling=1)
    jpg("base_420.jpg", photo, quality=80, subsampling=2)
    jpg("prog_444.jpg", photo, quality=90, subsampling=0, progressive=True)
    jpg("prog_420.jpg", photo, quality=75, subsampling=2, progressive=True)
    jpg("base_444_optimized.jpg", photo, quality=95, subsampling=0, optimize=True)
    jpg("base_420_restart.jpg", smooth(64, 128, 21), quality=80, subsampling=2, restart_marker_blocks=3)
    jpg("prog_444_restart.jpg", smooth(40, 72, 22), quality=85, subsampling=0, progressive=True, restart_marker_rows=1)
    jpg("grey.jpg", smooth(45, 70, 23)[..., 1], mode="L", quality=88)
    jpg("grey_prog.jpg", smooth(45, 70, 24)[..., 0], mode="L", quality=70, progressive=True)
    jpg("noise_444.jpg", noise(32, 48, 3, 25), quality=100, subsampling=0)          # saturating inverse DCT outputs
    jpg("narrow_420.jpg", smooth(19, 3, 26), quality=90, subsampling=2)              # chroma plane 2 samples wide: replication, not triangle
    jpg("one_pixel.jpg", smooth(1, 1, 27), quality=90, subsampling=0)

    def strip16(samples):
        """16-bit samples keep their high byte (libpng's png_set_strip_16, which DevIL calls); Pillow saturates 16-bit grey
        instead, so these expectations are computed from the source samples"""
        s8 = (samples >> 8).astype(np.uint32)
        if s8.ndim == 2:
            s8 = s8[..., None]
        c = s8.shape[-1]
        r, g, b = (s8[..., 0],) * 3 if c < 3 else (s8[..., 0], s8[..., 1], s8[..., 2])
        a = s8[..., 1] if c == 2 else (s8[..., 3] if c == 4 else np.uint32(255))
        bgra = (b | (g << np.uint32(8)) | (r << np.uint32(16)) | (a << np.uint32(24))).astype(np.uint32)
        gray = (s8[..., 0] if c < 3 else ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16)).astype(np.uint8)
        return bgra, gray

    sixteen = {"grey16.png": g16, "rgb16.png": (_hash(h * w * 3, 17) >> np.uint32(7)).astype(np.uint16).reshape(h, w, 3),
               "rgba16_adam7.png": (_hash(h * w * 4, 18) >> np.uint32(7)).astype(np.uint16).reshape(h, w, 4)}
    # a grey file whose only component declares 2x2 sampling factors (some encoders write that; T.81 A.2.2: a single
    # component is never interleaved, so the factors mean nothing and the image is the same)
    data = bytearray(open(os.path.join(HERE, "grey.jpg"), "rb").read())
    sof = data.index(b"\xff\xc0")
    assert data[sof + 9] == 1 and data[sof + 11] == 0x11
    data[sof + 11] = 0x22
    with open(os.path.join(HERE, "grey_h2v2.jpg"), "wb") as f:
        f.write(bytes(data))
    files["grey_h2v2.jpg"] = None

    arrays = {}
    for name in sorted(files):
        bgra, gray = strip16(sixteen[name]) if name in sixteen else expected(os.path.join(HERE, name))
        arrays[name + ":bgra"] = bgra
        arrays[name + ":l8"] = gray
    np.savez_compressed(os.path.join(HERE, "expected.npz"), **arrays)
    total = sum(os.path.getsize(os.path.join(HERE, n)) for n in os.listdir(HERE))
    print(f"{len(files)} files, {total} bytes in {HERE}")


if __name__ == "__main__":
    main()
