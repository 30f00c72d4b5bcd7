"""Builds libckd_b200.so in-tree: hand-written sm_100a CUDA kernels + the C ABI (include/ckd.h) + the C++ host layer.

    python -m cookiedough_b200.build

nvcc cross-compiles without a GPU.  Flags that matter for parity with the x86 reference:
-fmad=false (the reference has no FMA contraction), default -prec-div=true -prec-sqrt=true -ftz=false,
and -ffp-contract=off for the host-side per-frame set-up.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OUT = os.path.join(HERE, "libckd_b200.so")
OBJ = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall,-Wno-unused-function",
    "-Xptxas", "-v", *os.environ.get("CKD_EXTRA_NVCC", "").split(),
]


def sources():
    cu = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp")) if os.path.isdir(HOST) else []
    return cu, cpp


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = []
    for d in (CSRC, HOST, os.path.join(os.path.dirname(HERE), "include")):
        if os.path.isdir(d):
            hs += [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".h", ".cuh", ".hpp"))]
    return hs


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout)
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    cu, cpp = sources()
    headers = _headers()
    jobs, objs = [], []
    for src in cu:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append(([NVCC, *NVCC_FLAGS, "-c", src, "-o", obj], obj + ".log"))
    for src in cpp:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append((["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-msse4.1", "-Wall", "-c", src, "-o", obj], obj + ".log"))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        outs = list(ex.map(lambda j: _run(*j), jobs))
    if verbose:
        for o in outs:
            print(o)
    if jobs or not os.path.exists(OUT):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs, "-Xcompiler", "-fPIC", "-lz"], os.path.join(OBJ, "link.log"))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
