#!/usr/bin/env python
"""Per-SASS-instruction warp-state samples of one kernel from an ncu report captured with
--set full --import-source on:   tools/ncu_sass_samples.py report.ncu-rep [kernel-index] [min-samples]
Prints total stall-reason shares, the instructions holding >= min-samples samples with their top
reasons, and the execution count of every mbarrier try_wait (retry paths show up as extra
executions)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_s = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] and row:
        cur["rows"].append(row)
b = blocks[kidx]
h = b["hdr"]
col = {n: i for i, n in enumerate(h)}
reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
print(f"kernel[{kidx}] {b['name'][:110]}  ({len(blocks)} kernels in report)")
tot = {r: 0 for r in reasons}
nsmp = 0
for r in b["rows"]:
    nsmp += int(r[col["# Samples"]] or 0)
    for x in reasons:
        tot[x] += int(r[col[x]] or 0)
print(f"samples {nsmp}; " + ", ".join(f"{k[6:]} {100.0 * v / max(1, nsmp):.1f}%"
                                      for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
issued = sum(int(r[col["Instructions Executed"]] or 0) for r in b["rows"])
print(f"warp-instructions executed {issued:.3e}")
print("  idx    smp      exec  instruction / top reasons")
for i, r in enumerate(b["rows"]):
    smp = int(r[col["# Samples"]] or 0)
    src = r[col["Source"]].strip()
    if smp >= min_s or "TRYWAIT" in src or "UTCHMMA" in src[:12] or "STTM" in src:
        top = sorted(((int(r[col[x]] or 0), x[6:]) for x in reasons), reverse=True)[:3]
        tops = " ".join(f"{n}={v}" for v, n in top if v)
        print(f"{i:5d} {smp:6d} {int(r[col['Instructions Executed']] or 0):9d}  {src[:70]:70s} {tops}")
