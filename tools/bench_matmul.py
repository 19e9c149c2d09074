#!/usr/bin/env python3
"""Shape sweep over tools/bench_matmul, the role of the reference's
tools/benchmarks/matmul.py:92-195: the same default entries (M in {16, 256, 512} x the
Llama-3 8B/70B projection shapes), the same flags (--backend/--atype/--btype/--ctype),
`-algo tune` for every entry, stdout passed through.  Adds --csv to collect the best
solution per entry (the reference leaves parsing to the reader) and --table to write them
as a tuned-solution table: `PETIT_TUNE_TABLE=<file>` (or `petit_tune_table_load`) makes
the library's default chooser use it."""
import argparse
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))

SHAPES = [(4096, 4096), (4096, 14336), (6144, 4096), (8192, 8192), (8192, 28672), (10240, 8192),
          (28672, 4096), (57344, 8192)]
ENTRIES = [(m, n, k) for m in (16, 256, 512) for n, k in SHAPES]

WARMUP, REPEAT, BATCH, ALGO = 5, 20, 1, "tune"
LINE = re.compile(r"Matmul (\d+)x(\d+)x(\d+) .* algorithm: (\w*), (\d+) times total ([\d.]+) ms\. ([\d.]+) TFLOPS")


def run_benchmark(m, n, k, args):
    cmd = [os.path.join(ROOT, "bench_matmul"), "-backend", args.backend, "-atype", args.atype,
           "-btype", args.btype, "-ctype", args.ctype, "-m", str(m), "-k", str(k), "-n", str(n),
           "-warmup", str(WARMUP), "-repeat", str(REPEAT), "-batch", str(BATCH), "-algo", ALGO]
    if args.table:
        cmd += ["-table", args.table]
    try:
        result = subprocess.run(cmd, capture_output=True, text=True, timeout=args.timeout)
    except Exception as exc:  # noqa: BLE001
        print(f"Failed to run benchmark: {exc}", file=sys.stderr)
        return None
    print(result.stdout)
    if result.stderr:
        print(f"Error: {result.stderr}", file=sys.stderr)
    for line in result.stdout.splitlines():
        hit = LINE.match(line)
        if hit:  # first result line = fastest solution
            total_ms, tflops = float(hit.group(6)), float(hit.group(7))
            return {"m": m, "n": n, "k": k, "algo": hit.group(4), "us": total_ms * 1e3 / REPEAT,
                    "tflops": tflops}
    return None


def main():
    ap = argparse.ArgumentParser(description="Run matrix multiplication benchmarks")
    ap.add_argument("--backend", default="petit")
    ap.add_argument("--atype", default="fp16")
    ap.add_argument("--btype", default="nvfp4")
    ap.add_argument("--ctype", default="fp16")
    ap.add_argument("--csv", default=None, help="write the best solution per entry here")
    ap.add_argument("--table", default=None,
                    help="append '<btype> <atype> m n k <hex id>' of the fastest solution per entry")
    ap.add_argument("--timeout", type=float, default=300.0)
    args = ap.parse_args()
    rows = [r for r in (run_benchmark(m, n, k, args) for m, n, k in ENTRIES) if r]
    if args.csv:
        with open(args.csv, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=["m", "n", "k", "algo", "us", "tflops"])
            w.writeheader()
            w.writerows(rows)


if __name__ == "__main__":
    main()
