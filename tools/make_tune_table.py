"""Tuned-solution table from a `bench_matmul -algo tune` sweep log (tools/exp_midm.sh writes
"== n=N k=K m=M" followed by one result line per solution, fastest first).

  python tools/make_tune_table.py profiles/r02_mid_m_tune.log > petit-kernel_b200/petit_kernel/tuned/llama3_70b_b200.tune

Line format = what `bench_matmul -algo tune -table FILE` appends and PETIT_TUNE_TABLE /
petit_tune_table_load read: "<btype> <atype> m n k <solution id, 16 hex digits>"."""
import re
import sys


def main(path):
    out, key, best = [], None, None
    for line in open(path):
        mo = re.match(r"== n=(\d+) k=(\d+) m=(\d+)", line)
        if mo:
            if key and best:
                out.append((key, best))
            key, best = tuple(int(x) for x in mo.groups()), None
            continue
        mo = re.search(r"(\w+):(\w+)\. algorithm: ([0-9a-f]{16}), (\d+) times total ([\d.]+) ms", line)
        if mo and key:
            t = float(mo.group(5)) / int(mo.group(4))
            if best is None or t < best[0]:
                best = (t, mo.group(1), mo.group(3))
    if key and best:
        out.append((key, best))
    print("# Fastest token-tile variant per exact problem, NVFP4 x bf16, Llama-3.3-70B decoder GEMMs on one B200")
    print("# (measured with `bench_matmul -algo tune`, round 2; source log: " + path + ").")
    print("# Use: PETIT_TUNE_TABLE=<this file>, or petit_kernel.tuning.load_table(path).  Problems that are")
    print("# not listed fall back to the built-in rule (capi.cu default_ntok).")
    for (n, k, m), (t, atype, algo) in sorted(out, key=lambda e: (e[0][0], e[0][1], e[0][2])):
        print(f"nvfp4 {atype} {m} {n} {k} {algo}   # {t * 1e3:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
