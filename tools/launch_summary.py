#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) -> the per-kernel summary kept under profiles/.
    python tools/launch_summary.py gpurun_out/launches.csv "<the command that was profiled>" > profiles/rN_launches_summary.csv"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    rows.append((r[ik].replace("orlg::", ""), us))
tot = sum(u for _, u in rows)
agg = OrderedDict()
for k, u in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += u
print("# %s (every launch of the command; per-launch times under ncu are serialised and cold-cache: only the shares are meaningful)" % (
    sys.argv[2] if len(sys.argv) > 2 else "ncu --metrics gpu__time_duration.sum --clock-control none"))
print("kernel,launches,total_us,avg_us,share")
for k, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('"%s",%d,%.1f,%.2f,%.4f' % (k, c, u, u / c, u / tot))
ro = [(i, u) for i, (k, u) in enumerate(rows) if "deeprmsa_rollout_kernel<22, 0, 0, 2>" in k]
print("# rollout launches in order (index, us): " + " ".join("%d:%.0f" % x for x in ro[:40]))
