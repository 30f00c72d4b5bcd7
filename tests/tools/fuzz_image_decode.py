#!/usr/bin/env python3
"""ASAN/UBSAN fuzz of the host layer's PNG / JPEG decoders (cookiedough_b200/host/ckd_image.cpp).

Builds the decoder TU alone with -fsanitize=address,undefined next to a 20-line driver and feeds it seeded mutations of the
committed fixtures (byte flips with a bias to the headers, truncations, deletions; PNG chunk CRCs re-computed for most
mutants so that they reach the decoder body).  A mutant must decode or fail cleanly.

    python tests/tools/fuzz_image_decode.py [rounds=40]      # 300 mutants x 2 pixel formats per round
"""
import glob
import os
import random
import struct
import subprocess
import sys
import tempfile
import zlib

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DRIVER = r'''
#include <string>
#include <stdio.h>
static std::string g_err;
void SetLastError(const std::string &d) { g_err = d; }
extern "C" void *ckdhost_image_load(const char *path, int grayscale, int *width, int *height);
extern "C" void ckdhost_image_free(void *pixels);
int main(int argc, char **argv)
{
	int ok = 0, bad = 0;
	for (int i = 1; i < argc; ++i)
		for (int g = 0; g < 2; ++g)
		{
			int w = 0, h = 0;
			void *p = ckdhost_image_load(argv[i], g, &w, &h);
			if (p) { ++ok; ckdhost_image_free(p); } else ++bad;
		}
	printf("decoded %d failed %d\n", ok, bad);
	return 0;
}
'''


def fix_png_crcs(b):
    if b[:8] != b"\x89PNG\r\n\x1a\n":
        return b
    out, pos = bytearray(b[:8]), 8
    while pos + 12 <= len(b):
        ln = struct.unpack(">I", b[pos:pos + 4])[0]
        if ln > len(b) - pos - 12:
            break
        t, d = b[pos + 4:pos + 8], b[pos + 8:pos + 8 + ln]
        out += b[pos:pos + 8] + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
        pos += 12 + ln
    return bytes(out + b[pos:])


def mutate(data, rng, is_png):
    d = bytearray(data)
    kind = rng.random()
    if kind < 0.65:
        header = rng.random() < 0.5
        for _ in range(rng.choice((1, 1, 2, 4, 16))):
            d[rng.randrange(min(len(d), 700)) if header else rng.randrange(len(d))] = rng.randrange(256)
    elif kind < 0.8:
        d = d[:rng.randrange(1, len(d))]
    else:
        i = rng.randrange(len(d))
        d[i:i + rng.randrange(1, 64)] = b""
    d = bytes(d)
    return fix_png_crcs(d) if is_png and rng.random() < 0.7 else d


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    fixtures = sorted(glob.glob(os.path.join(REPO, "tests/golden/images/*.png")) + glob.glob(os.path.join(REPO, "tests/golden/images/*.jpg")))
    rng = random.Random(99)
    with tempfile.TemporaryDirectory(prefix="ckd_fuzz_") as tmp:
        with open(os.path.join(tmp, "main.cpp"), "w") as f:
            f.write(DRIVER)
        exe = os.path.join(tmp, "fuzz_decode")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-w", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                               "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "cookiedough_b200/host"),
                               os.path.join(tmp, "main.cpp"), os.path.join(REPO, "cookiedough_b200/host/ckd_image.cpp"), "-lz", "-o", exe])
        total = 0
        for rnd in range(rounds):
            files = []
            for path in fixtures:
                data = open(path, "rb").read()
                for m in range(10):
                    name = os.path.join(tmp, f"{rnd}_{m}_{os.path.basename(path)}")
                    with open(name, "wb") as f:
                        f.write(mutate(data, rng, path.endswith(".png")))
                    files.append(name)
            r = subprocess.run([exe] + files, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            total += len(files)
            if r.returncode != 0 or "ERROR" in r.stdout or "runtime error" in r.stdout:
                print(r.stdout[-4000:])
                sys.exit(f"round {rnd}: sanitizer report")
            for name in files:
                os.unlink(name)
        print(f"{total} mutants x 2 formats: clean")


if __name__ == "__main__":
    main()
