#!/usr/bin/env python3
"""oracle/build_ref.py -- TEST INFRASTRUCTURE: build the *reference itself* as a CPU oracle.

Compiles the reference's own hot-path translation units from where they lie under
/root/reference (nothing is copied into the repository) together with oracle/ref_shim.cpp into

    oracle/_ref/libckd_ref_720.so     (kResX x kResY = 1280 x 720, the reference as shipped)
    oracle/_ref/libckd_ref_2160.so    (3840 x 2160, resolution constants patched)

and prepares the reference's input data in refdata/ at the repository root (git-ignored, but shipped to the GPU box; these
are inputs -- timeline and art -- that the product reads as well, so they do not live under oracle/):

    refdata/sync/*.track            binary GNU Rocket tracks (target/sync)
    refdata/directors-cut.rocket    the XML Rocket project (target/directors-cut.rocket)
    refdata/assets.npz              art/maps decoded once with Pillow (BGRA / L8), shared by both sides

Patch list (applied on the fly to a throw-away symlink tree in a temp dir):
    P1  fx-blitter.cpp:48-49   _mm_load_si128 on a 4-byte aligned address -> _mm_loadu_si128 (x86 #GP)
    P2  polar.cpp:113,161      clamp tile rows to kResY (720 % 32 != 0 -> heap overflow); kFxMapResY for the FX-map instantiation
    P2b polar.cpp:118          Polar_Blit_Tile: clamp tile columns to xRes (Polar_Blit_2x2: 644 % 64 != 0; no-op at full size)
    P3a boxblur.cpp:18         kMaxRes 2048 -> 4096 (4K scratch)
    P3b shadertoy.cpp:185      blur-map scratch (1280*720)/2 px -> kFxMapBytes
    P3c main.h:37-38           kResX/kResY (4K build only)
    P3d demo.cpp:26            static_assert(kResX == 1280 && kResY == 720) dropped (4K build only)
The reference's CMake build is not used (it needs SDL2/DevIL/BASS); flags per SURVEY 8c:
    g++ -std=c++20 -O3 -msse4.1 -fopenmp -fno-exceptions -DSYNC_PLAYER ; gcc -O3 -DSYNC_PLAYER for Rocket's C.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

REF = os.environ.get("CKD_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
DATA = os.path.join(os.path.dirname(HERE), "refdata")   # decoded inputs of the reference: used by the product too, so not under oracle/

CPP_UNITS = [
    "shadertoy.cpp", "landscape.cpp", "tunnelscape.cpp", "ball.cpp", "torus-twister.cpp",
    "polar.cpp", "boxblur.cpp", "deprecated/boxblur.cpp", "fx-blitter.cpp", "util.cpp",
    "shared-resources.cpp", "sincos-lut.cpp", "fast-cosine.cpp", "rocket.cpp", "demo.cpp",
]
C_UNITS = ["track.c", "device.c"]  # 3rdparty/rocket-stripped/lib

CXXFLAGS = ["-std=c++20", "-O3", "-msse4.1", "-fopenmp", "-fno-exceptions", "-DSYNC_PLAYER", "-w", "-fPIC"]
CFLAGS = ["-O3", "-DSYNC_PLAYER", "-w", "-fPIC"]

# images the five X_Create() calls, Shared_Create() and Demo_Create() load: the list lives with the product's asset rules
sys.path.insert(0, os.path.dirname(HERE))
from cookiedough_b200.assets import SPEC as _SPEC  # noqa: E402

MISSING = {"assets/scape/tscape-C7W-edit.png"}  # listed in the reference's .MISSING_LARGE_BLOBS: synthesised from C17W-edit
ASSETS = [(path, spec[2]) for path, spec in _SPEC.items() if path not in MISSING]


def _patch(text, pattern, repl, count, what):
    new, n = re.subn(pattern, repl, text)
    if n != count:
        raise RuntimeError(f"patch '{what}' applied {n} times, expected {count}")
    return new


def make_tree(tmp, res_x, res_y):
    """symlink tree of code/ (+ real copies of the few patched files) so that every TU sees the patched main.h"""
    code = os.path.join(tmp, "code")
    os.makedirs(os.path.join(code, "deprecated"))
    os.symlink(os.path.join(REF, "3rdparty"), os.path.join(tmp, "3rdparty"))
    src = os.path.join(REF, "code")
    for name in os.listdir(src):
        p = os.path.join(src, name)
        if name == "deprecated":
            for sub in os.listdir(p):
                os.symlink(os.path.join(p, sub), os.path.join(code, "deprecated", sub))
        else:
            os.symlink(p, os.path.join(code, name))

    def rewrite(rel, fn):
        dst = os.path.join(code, rel)
        with open(os.path.join(src, rel), "r", encoding="utf-8", errors="replace") as f:
            text = f.read()
        os.unlink(dst)
        with open(dst, "w", encoding="utf-8") as f:
            f.write(fn(text))

    # P1
    rewrite("fx-blitter.cpp", lambda t: _patch(
        t, r"_mm_load_si128\(reinterpret_cast<const __m128i\*>\(&pSrc\[", "_mm_loadu_si128(reinterpret_cast<const __m128i*>(&pSrc[", 2, "P1"))
    # P2
    def p2(t):
        # first row loop = Polar_Blit_Tile<xRes> (also instantiated for the FX map by Polar_Blit_2x2), second = Polar_Blit_TileA
        row_loop = "for (unsigned iY = tY; iY < tY + tileSize; ++iY)"
        if t.count(row_loop) != 2:
            raise RuntimeError("patch 'P2': expected 2 tile row loops")
        t = t.replace(row_loop, "for (unsigned iY = tY; iY < tY + tileSize && iY < (xRes == kFxMapResX ? kFxMapResY : kResY); ++iY)", 1)
        t = t.replace(row_loop, "for (unsigned iY = tY; iY < tY + tileSize && iY < kResY; ++iY)", 1)
        # P2b: only Polar_Blit_Tile<xRes> steps its columns by 4
        return _patch(t, r"for \(unsigned iX = 0; iX < tileSize; iX \+= 4\)", "for (unsigned iX = 0; iX < tileSize && tX + iX < xRes; iX += 4)", 1, "P2b")
    rewrite("polar.cpp", p2)
    # P3a, P3b (harmless at 720p; applied to both builds so the two oracles share one code base)
    rewrite("boxblur.cpp", lambda t: _patch(t, r"constexpr size_t kMaxRes = 2048;", "constexpr size_t kMaxRes = 4096;", 1, "P3a"))
    rewrite("shadertoy.cpp", lambda t: _patch(
        t, r"mallocAligned\(\(1280\*720\)/2 \* sizeof\(uint32_t\), kAlignTo\)", "mallocAligned(kFxMapBytes, kAlignTo)", 1, "P3b"))
    # P3c
    if (res_x, res_y) != (1280, 720):
        def p3c(t):
            t = _patch(t, r"constexpr size_t kResX = 1280;", f"constexpr size_t kResX = {res_x};", 1, "P3c-x")
            return _patch(t, r"constexpr size_t kResY = 720;", f"constexpr size_t kResY = {res_y};", 1, "P3c-y")
        rewrite("main.h", p3c)
        rewrite("demo.cpp", lambda t: _patch(t, r"static_assert\(kResX == 1280 && kResY == 720\);", "", 1, "P3d"))
    return code


def run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + r.stdout)


def build_lib(res_x, res_y):
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, f"libckd_ref_{res_y}.so")
    with tempfile.TemporaryDirectory(prefix="ckd_ref_") as tmp:
        code = make_tree(tmp, res_x, res_y)
        jobs = []
        objs = []
        for unit in CPP_UNITS:
            obj = os.path.join(tmp, unit.replace("/", "_") + ".o")
            objs.append(obj)
            jobs.append(["g++", *CXXFLAGS, "-c", os.path.join(code, unit), "-o", obj])
        for unit in C_UNITS:
            obj = os.path.join(tmp, unit + ".o")
            objs.append(obj)
            jobs.append(["gcc", *CFLAGS, "-c", os.path.join(tmp, "3rdparty/rocket-stripped/lib", unit), "-o", obj])
        shim_obj = os.path.join(tmp, "ref_shim.o")
        objs.append(shim_obj)
        jobs.append(["g++", *CXXFLAGS, "-I", code, "-iquote", code, "-c", os.path.join(HERE, "ref_shim.cpp"), "-o", shim_obj])
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            list(ex.map(run, jobs))
        run(["g++", "-shared", "-fopenmp", "-o", lib, *objs, "-lm"])
    return lib


def prepare_data():
    data = DATA
    sync = os.path.join(data, "sync")
    os.makedirs(sync, exist_ok=True)
    src_sync = os.path.join(REF, "target", "sync")
    for name in os.listdir(src_sync):
        if name.endswith(".track"):
            shutil.copyfile(os.path.join(src_sync, name), os.path.join(sync, name))
    shutil.copyfile(os.path.join(REF, "target", "directors-cut.rocket"), os.path.join(data, "directors-cut.rocket"))


def prepare_assets():
    import numpy as np
    from PIL import Image

    arrays = {}
    for path, is_gray in ASSETS:
        img = Image.open(os.path.join(REF, "target", path))
        if is_gray:
            arr = np.asarray(img.convert("L"), dtype=np.uint8)
        else:
            rgba = np.asarray(img.convert("RGBA"), dtype=np.uint8)
            arr = np.ascontiguousarray(rgba[..., [2, 1, 0, 3]])  # -> B,G,R,A bytes == little-endian 0xAARRGGBB (code/image.cpp:53-54)
            arr = arr.view(np.uint32).reshape(arr.shape[0], arr.shape[1])
        arrays[path] = arr
    os.makedirs(DATA, exist_ok=True)
    np.savez_compressed(os.path.join(DATA, "assets.npz"), **arrays)


def main():
    if not os.path.isdir(os.path.join(REF, "code")):
        print(f"build_ref: reference not found at {REF}; keeping prebuilt oracle/_ref (if any)")
        return 0
    libs = [build_lib(1280, 720), build_lib(3840, 2160)]
    prepare_data()
    prepare_assets()
    for lib in libs:
        print("built", lib)
    return 0


if __name__ == "__main__":
    sys.exit(main())
