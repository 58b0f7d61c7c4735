#!/usr/bin/env python
"""Instruction count and stall samples per source-line RANGE (phases) of the fast kernel.
    python tools/ncu_phases.py <rep> <mangled-regex>
"""
import sys
from collections import defaultdict
sys.path.insert(0, 'tools')
import ncu_lines as N

rep, mre = sys.argv[1], sys.argv[2]
name, hdr, data = N.sass_rows(rep, 'deeprmsa_fast_kernel')
ix = {h: i for i, h in enumerate(hdr)}
table = N.line_table(mre)
print(len(data), len(table))
src = open('optical-rl-gym_b200/csrc/orlg_deeprmsa_fast.cuh').read().splitlines()
# phase boundaries by PHASE_MARK lines in the fast kernel file
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if 'PHASE_MARK(' in l and '#define' not in l and 'do {' not in l]
def phase_of(file, line):
    if file != 'orlg_deeprmsa_fast.cuh':
        return None
    k = 0
    for j, (ln, _) in enumerate(marks):
        if line > ln:
            k = j + 1
    return k
names = ["issue loads", "wait tables", "A decision+push", "B traffic", "wait masks", "alloc+release", "path AND", "writeback", "barrier1", "features", "scalar stores", "barrier2", "copy-out", "end"]
agg = defaultdict(lambda: [0, 0])
last_phase = 0
tot_i = tot_s = 0
for i in range(min(len(data), len(table))):
    f, ln = table[i][0], table[i][1]
    ph = phase_of(f, ln)
    if ph is None:
        ph = last_phase          # inlined helper: attribute to the surrounding phase
    else:
        last_phase = ph
    ie = int(data[i][ix['Instructions Executed']]); sm = int(data[i][ix['# Samples']])
    agg[ph][0] += ie; agg[ph][1] += sm; tot_i += ie; tot_s += sm
for ph in sorted(agg):
    print("%-18s inst/warp %7.1f (%4.1f%%)  samples %5d (%4.1f%%)" % (names[ph] if ph < len(names) else ph, agg[ph][0] / 2048, 100 * agg[ph][0] / tot_i, agg[ph][1], 100 * agg[ph][1] / tot_s))
print("total inst/warp %.0f samples %d" % (tot_i / 2048, tot_s))
