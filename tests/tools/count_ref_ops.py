#!/usr/bin/env python3
"""tests/tools/count_ref_ops.py -- TEST INFRASTRUCTURE: the op-counting build of the REFERENCE.

The roofline of the raymarch kernels needs the algorithmic FP work per FX-map pixel, and early exits make that data
dependent.  This tool measures it on the reference itself, independently of this repository's kernels:

  1. the reference's shadertoy.cpp (with everything it inlines: shadertoy-util.h, sincos-lut.h, Std3DMath, sse_mathfun.h) is
     compiled to assembly with the oracle's own flags (oracle/build_ref.py, + -mno-red-zone),
  2. every SSE floating-point instruction in that assembly gets a flag-preserving counter increment in front of it
     (pushfq / addq $lanes, counter(%rip) / popfq) -- an instruction-level "counting float",
  3. the instrumented unit is linked with the other, untouched reference units into oracle/_ref/libckd_ref_2160_count.so,
  4. every raymarch effect of bench.py's suite is drawn once at its pinned Rocket row at 3840x2160 on ONE thread through the
     reference's own X_Draw, and the tallies are divided by the FX-map pixel count.

Counting rule: 1 op = one FP add / sub / mul / div / sqrt / rsqrt / compare / min-max / conversion (round
included) per LANE: a packed instruction counts 4 (cvtpd2ps 2), dpps counts 4 mul + 3 add; a libm call (powf, expf, atan2f ...)
counts as ONE op (a floor; the calls are tallied separately).  Bit logic on float registers (andps / xorps / blendvps ...),
moves and shuffles are tallied but are not ops.  The frame's output is checked against the uninstrumented oracle, so the
instrumentation provably did not change the computation.

    python tests/tools/count_ref_ops.py            -> profiles/r02_ref_fp_ops.json   (needs /root/reference; CPU only)
"""
import collections
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

COUNTERS = ["fadd", "fmul", "fdiv", "fsqrt", "frsqrt", "fcmp", "fminmax", "fcvt", "libm_calls", "logic", "instr_scalar", "instr_packed"]
IDX = {n: i for i, n in enumerate(COUNTERS)}
LIBM = {"powf", "expf", "logf", "sinf", "cosf", "tanf", "tanhf", "atan2f", "atanf", "acosf", "asinf", "fmodf", "roundf", "floorf", "ceilf",
        "pow", "exp", "log", "sin", "cos", "atan2", "sqrtf"}


def classify(mn):
    """-> list of (counter, amount) for one mnemonic, or None when it is not a floating-point operation"""
    packed = mn.endswith(("ps", "pd")) and not mn.startswith("cvt")
    lanes = 4 if mn.endswith("ps") else 2 if mn.endswith("pd") else 1
    if re.fullmatch(r"(add|sub|addsub|hadd|hsub)(ss|ps|sd|pd)", mn):
        return [("fadd", lanes)], packed
    if re.fullmatch(r"mul(ss|ps|sd|pd)", mn):
        return [("fmul", lanes)], packed
    if re.fullmatch(r"div(ss|ps|sd|pd)", mn):
        return [("fdiv", lanes)], packed
    if re.fullmatch(r"sqrt(ss|ps|sd|pd)", mn):
        return [("fsqrt", lanes)], packed
    if re.fullmatch(r"(rsqrt|rcp)(ss|ps)", mn):
        return [("frsqrt", lanes)], packed
    if mn == "dpps":
        return [("fmul", 4), ("fadd", 3)], True
    if re.fullmatch(r"u?comis[sd]", mn) or re.fullmatch(r"cmp[a-z]*(ss|sd)", mn):
        return [("fcmp", 1)], False
    if re.fullmatch(r"cmp[a-z]*(ps|pd)", mn):
        return [("fcmp", lanes)], True
    if re.fullmatch(r"(min|max)(ss|ps|sd|pd)", mn):
        return [("fminmax", lanes)], packed
    if re.fullmatch(r"round(ss|ps|sd|pd)", mn):
        return [("fcvt", lanes)], packed
    if mn.startswith("cvt"):
        if re.fullmatch(r"cvtt?ps2dq|cvtdq2ps", mn):
            return [("fcvt", 4)], True
        if re.fullmatch(r"cvtpd2ps|cvtps2pd|cvtt?pd2dq|cvtdq2pd", mn):
            return [("fcvt", 2)], True
        return [("fcvt", 1)], False           # cvt(t)ss2si, cvtsi2ss, cvtss2sd, cvtsd2ss (+ l/q suffixes)
    if re.fullmatch(r"(and|andn|or|xor)(ps|pd)|blendv?(ps|pd)", mn):
        return [("logic", lanes)], None
    return None


def instrument(asm_in, asm_out):
    """writes the instrumented assembly; returns the static histogram of counted mnemonics"""
    hist = collections.Counter()
    out = ["\t.hidden ckd_ref_ops\n"]

    def bump(counter, amount):
        return f"\taddq\t${amount}, ckd_ref_ops+{8 * IDX[counter]}(%rip)\n"
    with open(asm_in) as f:
        for line in f:
            m = re.match(r"\s+([a-z][a-z0-9]*)\s", line)
            if m:
                mn = m.group(1)
                incs = []
                if mn in ("call", "jmp"):
                    t = re.match(r"\s+(?:call|jmp)\s+([A-Za-z_][A-Za-z0-9_]*)(@PLT)?\s*$", line)
                    if t and t.group(1) in LIBM:
                        incs = [("libm_calls", 1)]
                        hist["call " + t.group(1)] += 1
                else:
                    c = classify(mn)
                    if c is not None:
                        incs, packed = list(c[0]), c[1]
                        if packed is not None:
                            incs.append(("instr_packed" if packed else "instr_scalar", 1))
                        hist[mn] += 1
                if incs:
                    out.append("\tpushfq\n")
                    out.extend(bump(cn, amt) for cn, amt in incs)
                    out.append("\tpopfq\n")
            out.append(line)
    with open(asm_out, "w") as f:
        f.writelines(out)
    return hist


def build_counting_lib(res_x=3840, res_y=2160):
    import build_ref as B
    os.makedirs(B.OUT, exist_ok=True)
    lib = os.path.join(B.OUT, f"libckd_ref_{res_y}_count.so")
    with tempfile.TemporaryDirectory(prefix="ckd_cnt_") as tmp:
        code = B.make_tree(tmp, res_x, res_y)
        jobs, objs = [], []
        for unit in B.CPP_UNITS:
            if unit == "shadertoy.cpp":
                continue
            obj = os.path.join(tmp, unit.replace("/", "_") + ".o")
            objs.append(obj)
            jobs.append(["g++", *B.CXXFLAGS, "-c", os.path.join(code, unit), "-o", obj])
        for unit in B.C_UNITS:
            obj = os.path.join(tmp, unit + ".o")
            objs.append(obj)
            jobs.append(["gcc", *B.CFLAGS, "-c", os.path.join(tmp, "3rdparty/rocket-stripped/lib", unit), "-o", obj])
        shim_obj = os.path.join(tmp, "ref_shim.o")
        objs.append(shim_obj)
        jobs.append(["g++", *B.CXXFLAGS, "-I", code, "-iquote", code, "-c", os.path.join(B.HERE, "ref_shim.cpp"), "-o", shim_obj])
        asm = os.path.join(tmp, "shadertoy.s")
        jobs.append(["g++", *B.CXXFLAGS, "-mno-red-zone", "-S", os.path.join(code, "shadertoy.cpp"), "-o", asm])
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            list(ex.map(B.run, jobs))
        asm_i = os.path.join(tmp, "shadertoy_counted.s")
        hist = instrument(asm, asm_i)
        obj = os.path.join(tmp, "shadertoy_counted.o")
        B.run(["g++", "-c", asm_i, "-o", obj])
        objs.append(obj)
        cnt_c = os.path.join(tmp, "counters.c")
        with open(cnt_c, "w") as f:
            f.write(f"__attribute__((visibility(\"hidden\"))) unsigned long long ckd_ref_ops[{len(COUNTERS)}];\n"
                    "unsigned long long *ckd_ref_ops_ptr(void) { return ckd_ref_ops; }\n")
        cnt_o = os.path.join(tmp, "counters.o")
        B.run(["gcc", "-O2", "-fPIC", "-c", cnt_c, "-o", cnt_o])
        objs.append(cnt_o)
        B.run(["g++", "-shared", "-fopenmp", "-o", lib, *objs, "-lm"])
    return lib, hist


# bench.py name -> (reference effect, pinned row)
KERNELS = {
    "raymarch_plasma": ("plasma", 2600), "raymarch_nautilus": ("nautilus", 5700), "raymarch_spikey_close": ("spikey_close", 6800),
    "raymarch_spikey_distant": ("spikey_distant", 3600), "raymarch_tunnel": ("tunnel", 4500), "raymarch_sinuses": ("sinuses", 7800),
    "raymarch_laura": ("laura", 8900),
}


def main():
    import numpy as np
    res_x, res_y = 3840, 2160
    os.environ["OMP_NUM_THREADS"] = "1"      # the counters are plain globals: one thread
    lib, hist = build_counting_lib(res_x, res_y)
    from oracle import ref as oref
    from cookiedough_b200.assets import Assets
    plain = oref.lib_path
    oref.lib_path = lambda ry: lib if ry == res_y else plain(ry)
    R = oref.Reference(res_y, Assets(res_x, res_y))
    R.lib.ckd_ref_ops_ptr.restype = C.POINTER(C.c_ulonglong)
    ops = R.lib.ckd_ref_ops_ptr()
    fx_pixels = R.fx_x * R.fx_y

    # the uninstrumented oracle in a child process (the reference keeps its state in globals: one library per process)
    def plain_sha(effect, row):
        code = ("import sys, hashlib; sys.path.insert(0, %r); from oracle import ref as o; from cookiedough_b200.assets import Assets;"
                "R = o.Reference(%d, Assets(%d, %d)); R.set_row(%d); print(hashlib.sha256(R.draw(%r).tobytes()).hexdigest())"
                % (REPO, res_y, res_x, res_y, row, effect))
        return subprocess.check_output([sys.executable, "-c", code], text=True).strip().splitlines()[-1]

    import hashlib
    out = {}
    frame = R.frame()
    for name, (effect, row) in KERNELS.items():
        R.set_row(row)
        for i in range(len(COUNTERS)):
            ops[i] = 0
        R.draw(effect, frame)
        tally = {n: int(ops[i]) for i, n in enumerate(COUNTERS)}
        same = hashlib.sha256(frame.tobytes()).hexdigest() == plain_sha(effect, row)
        total = sum(tally[n] for n in ("fadd", "fmul", "fdiv", "fsqrt", "frsqrt", "fcmp", "fminmax", "fcvt", "libm_calls"))
        rec = {n: tally[n] / fx_pixels for n in COUNTERS}
        rec.update({"ops_per_fx_pixel": total / fx_pixels, "effect": effect, "row": row, "frame_identical_to_plain_oracle": bool(same)})
        out[name] = rec
        print(f"{name:26s} {rec['ops_per_fx_pixel']:8.1f} ops/FX px  (fadd {rec['fadd']:.0f} fmul {rec['fmul']:.0f} cvt {rec['fcvt']:.0f} cmp {rec['fcmp']:.0f} "
              f"minmax {rec['fminmax']:.0f} div {rec['fdiv']:.1f} sqrt {rec['fsqrt']:.1f} rsqrt {rec['frsqrt']:.1f} libm {rec['libm_calls']:.2f}; "
              f"scalar instr {rec['instr_scalar']:.0f}, packed instr {rec['instr_packed']:.0f}; logic {rec['logic']:.0f})  identical={same}", flush=True)
        assert same, "the instrumented build changed the frame"
    doc = {"res": [res_x, res_y], "fx_pixels": fx_pixels, "source": "instrumented build of the reference's shadertoy.cpp (tests/tools/count_ref_ops.py)",
           "rule": "1 op = one FP add/sub/mul/div/sqrt/rsqrt/compare/min-max/conversion per lane (packed x4, dpps = 4 mul + 3 add); one libm call = 1 op; "
                   "float-register bit logic, moves and shuffles are not ops",
           "flags": "g++ -std=c++20 -O3 -msse4.1 -fopenmp -fno-exceptions -DSYNC_PLAYER -mno-red-zone (oracle/build_ref.py flags)",
           "static_histogram": dict(hist.most_common()), "kernels": out}
    path = os.path.join(REPO, "profiles", "r02_ref_fp_ops.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
