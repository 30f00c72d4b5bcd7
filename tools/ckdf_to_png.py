#!/usr/bin/env python3
"""Export frames of a CKDF stream (tools/render_demo.py, host/ckd_sink.cpp) as PNG files.

    python tools/ckdf_to_png.py demo.ckdf out_dir [first [last [step]]]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookiedough_b200 import sink


def main():
    from PIL import Image
    path, out_dir = sys.argv[1], sys.argv[2]
    res_x, res_y, n = sink.read_header(path)
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    last = int(sys.argv[4]) if len(sys.argv) > 4 else n - 1
    step = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    os.makedirs(out_dir, exist_ok=True)
    for i in range(first, min(last, n - 1) + 1, step):
        bgra = sink.read_frame(path, i).view(np.uint8).reshape(res_y, res_x, 4)
        rgb = np.ascontiguousarray(bgra[..., [2, 1, 0]])   # 0xAARRGGBB little-endian = B, G, R, A bytes; the demo's alpha is not coverage
        Image.fromarray(rgb, "RGB").save(os.path.join(out_dir, f"frame_{i:05d}.png"))
    print(f"{path}: {res_x}x{res_y}, {n} frames; wrote frames {first}..{min(last, n - 1)} step {step} to {out_dir}")


if __name__ == "__main__":
    main()
