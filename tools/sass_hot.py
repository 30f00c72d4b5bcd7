#!/usr/bin/env python3
"""reads an `ncu --page source --print-source sass --csv` export (optionally .gz) and prints the stall picture:
totals per stall reason, and the hottest instructions with their dominant stall reasons.
    python tools/sass_hot.py gpurun_out/ncu/x_sass.csv.gz [top N] [from-addr-index to-addr-index]"""
import csv, gzip, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
print(rows[0][1][:150])
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = rows[2:]
tot_samples = sum(int(r[col["# Samples"]]) for r in data)
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in data)
print(f"instructions (static) {len(data)}, warp instructions executed {tot_inst}, samples {tot_samples}")
tot = {n: sum(int(r[col[n]]) for r in data) for n in stall_cols}
print("stall totals:", ", ".join(f"{n[6:]} {v} ({100.0*v/max(1,tot_samples):.1f}%)" for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]]))[:top]
print("hottest instructions (index, samples, executed, source, main stalls):")
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[col[n]]), n[6:]) for n in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r[col['# Samples']]):7d} {int(r[col['Instructions Executed']]):9d}  {r[col['Source']].strip():60s} " + " ".join(f"{n}:{v}" for v, n in st if v))
if len(sys.argv) > 4:
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for i in range(a, b):
        r = data[i]
        st = sorted(((int(r[col[n]]), n[6:]) for n in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {int(r[col['# Samples']]):7d} {int(r[col['Instructions Executed']]):9d}  {r[col['Source']].strip():60s} " + " ".join(f"{n}:{v}" for v, n in st if v))
