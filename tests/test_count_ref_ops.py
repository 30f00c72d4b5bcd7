"""The instruction classifier / instrumenter of the reference op counter (tests/tools/count_ref_ops.py): which SSE mnemonics count
as which op, with how many lanes, and what the instrumented assembly looks like.  (The counting build itself needs the reference
checkout and a few minutes: it is run by hand and its result is committed as profiles/r02_ref_fp_ops.json.)"""
import json
import os
import sys

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "tests", "tools"))
import count_ref_ops as cro  # noqa: E402


def test_classification_of_sse_mnemonics():
    assert cro.classify("addss") == ([("fadd", 1)], False)
    assert cro.classify("subps") == ([("fadd", 4)], True)
    assert cro.classify("mulps") == ([("fmul", 4)], True)
    assert cro.classify("divss") == ([("fdiv", 1)], False)
    assert cro.classify("sqrtss") == ([("fsqrt", 1)], False)
    assert cro.classify("rsqrtps") == ([("frsqrt", 4)], True)
    assert cro.classify("dpps") == ([("fmul", 4), ("fadd", 3)], True)
    assert cro.classify("comiss") == ([("fcmp", 1)], False)
    assert cro.classify("cmpltps") == ([("fcmp", 4)], True)
    assert cro.classify("maxss") == ([("fminmax", 1)], False)
    assert cro.classify("roundss") == ([("fcvt", 1)], False)
    assert cro.classify("cvttss2sil") == ([("fcvt", 1)], False)
    assert cro.classify("cvtps2dq") == ([("fcvt", 4)], True)
    assert cro.classify("cvtpd2ps") == ([("fcvt", 2)], True)
    assert cro.classify("andps")[0] == [("logic", 4)]
    for not_an_op in ("movss", "movaps", "shufps", "unpcklps", "pmaxsd", "paddd", "leaq", "call"):
        assert cro.classify(not_an_op) is None, not_an_op


def test_instrumented_assembly(tmp_path):
    src = tmp_path / "in.s"
    src.write_text("\t.text\nf:\n\tmovss\t(%rdi), %xmm0\n\tmulss\t%xmm1, %xmm0\n\taddps\t%xmm2, %xmm0\n\tcall\tpowf@PLT\n\tcall\tother@PLT\n\tret\n")
    out = tmp_path / "out.s"
    hist = cro.instrument(str(src), str(out))
    text = out.read_text()
    assert hist == {"mulss": 1, "addps": 1, "call powf": 1}
    # every counted instruction is preceded by a flag-preserving bump of its counters; nothing else is touched
    assert text.count("pushfq") == 3 and text.count("popfq") == 3
    assert f"addq\t$1, ckd_ref_ops+{8 * cro.IDX['fmul']}(%rip)" in text
    assert f"addq\t$4, ckd_ref_ops+{8 * cro.IDX['fadd']}(%rip)" in text
    assert f"addq\t$1, ckd_ref_ops+{8 * cro.IDX['libm_calls']}(%rip)" in text
    assert f"addq\t$1, ckd_ref_ops+{8 * cro.IDX['instr_packed']}(%rip)" in text
    lines = [l for l in text.splitlines() if "pushfq" not in l and "popfq" not in l and "ckd_ref_ops" not in l]
    assert lines == src.read_text().splitlines()


def test_committed_counts_are_consistent():
    with open(os.path.join(REPO, "profiles", "r02_ref_fp_ops.json")) as f:
        doc = json.load(f)
    assert doc["fx_pixels"] == (3840 // 2 + 4) * (2160 // 2 + 4)
    for name, k in doc["kernels"].items():
        total = sum(k[c] for c in ("fadd", "fmul", "fdiv", "fsqrt", "frsqrt", "fcmp", "fminmax", "fcvt", "libm_calls"))
        assert abs(total - k["ops_per_fx_pixel"]) < 1e-6, name
        assert k["frame_identical_to_plain_oracle"] is True, name


def test_bench_reads_the_reference_counted_denominators():
    sys.path.insert(0, REPO)
    import bench
    ops, addmul, source = bench.flop_table()
    assert "counted on the reference" in source
    with open(os.path.join(REPO, "profiles", "r02_ref_fp_ops.json")) as f:
        doc = json.load(f)
    for name in bench.FLOP_FALLBACK:
        assert ops[name] == doc["kernels"][name]["ops_per_fx_pixel"]
        assert addmul[name] == doc["kernels"][name]["fadd"] + doc["kernels"][name]["fmul"]
    # both arms of the bench describe the same workload
    assert bench.suite_config() == {"workload": "effect-suite-4k", "res": [3840, 2160], "effects": [s[0] for s in bench.SUITE]}


def test_bench_stdout_carries_only_the_result_line():
    """bench.py owes the driver exactly one JSON line on stdout; whatever else the process or a library writes to fd 1 (NCCL's
    version banner at N > 1) must come out on stderr"""
    import json
    import subprocess
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version x.y\\n'); print('chatter'); bench.emit({'value': 1})") % REPO
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"value": 1}
    assert "NCCL version x.y" in r.stderr and "chatter" in r.stderr
