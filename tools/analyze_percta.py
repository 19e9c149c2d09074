#!/usr/bin/env python
"""Summarise the per-CTA two-launch traces written by `PETIT_TRACE2=1 PETIT_TRACE_DUMP=f
tools/gemm_bench ...`: per-event percentiles relative to the moment griddepcontrol.wait
returns in launch 1, the main-loop duration per CTA and the last CTAs to exit."""
import collections
import statistics as st
import sys


def load(fn):
    data = collections.defaultdict(list)
    names = None
    for line in open(fn):
        if line.startswith('#'):
            names = line.strip().split(': ')[1].split(',')
            continue
        f = line.strip().split(',')
        data[(f[0], int(f[2]))].append([int(f[3]), int(f[4])] + [float(x) for x in f[5:]])
    return data, names


for fn in sys.argv[1:]:
    data, names = load(fn)
    ix = {n: i + 2 for i, n in enumerate(names[5:])}
    for (shape, launch), rows in data.items():
        if launch != 1:
            continue
        rows.sort()

        def col(nm):
            return [r[ix[nm]] for r in rows]

        t_wait = min(col('griddep_wait_done'))
        prev_exit = max(r[ix['exit']] for r in data[(shape, 0)])
        print(f"== {shape} launch 1, {len(rows)} CTAs; us relative to the first griddep_wait return "
              f"(previous launch's last exit at {prev_exit - t_wait:+.2f})")
        for nm in ['entry', 'setup_done', 'griddep_wait_done', 'first_stage_landed', 'dequant_done',
                   'mma_issued_all', 'last_acc_full', 'epilogue_done', 'exit']:
            c = sorted(x - t_wait for x in col(nm) if x)
            if c:
                print(f"   {nm:20s} min {c[0]:6.2f}  p50 {st.median(c):6.2f}  p90 {c[int(.9 * len(c))]:6.2f}"
                      f"  max {c[-1]:6.2f}")
        d = sorted(r[ix['dequant_done']] - r[ix['first_stage_landed']] for r in rows)
        print(f"   main loop per CTA    min {d[0]:6.2f}  p50 {st.median(d):6.2f}  p90 {d[int(.9 * len(d))]:6.2f}"
              f"  max {d[-1]:6.2f}")
        for r in sorted(rows, key=lambda r: -r[ix['exit']])[:6]:
            g = lambda nm: (r[ix[nm]] - t_wait) if r[ix[nm]] else float('nan')
            print(f"   late cta {r[0]:3d} sm {r[1]:3d}: entry {g('entry'):6.2f} wait {g('griddep_wait_done'):6.2f}"
                  f" first {g('first_stage_landed'):6.2f} dq_done {g('dequant_done'):6.2f}"
                  f" acc_full {g('last_acc_full'):6.2f} polled {g('lastseg_polled'):6.2f} exit {g('exit'):6.2f}")
