#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-substring> [top_n] [mangled-regex]

Joins `ncu --page source --csv --print-source sass` (per-SASS-instruction samples) with the line
table of the in-tree liborlg.so (`cuobjdump -xelf` + `nvdisasm -g`), matching instructions by their
order inside the kernel.  Needs the library built with -lineinfo (it is).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "optical-rl-gym_b200", "optical_rl_gym_b200", "liborlg.so")


def sass_rows(rep, kernel_sub):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, cur, hdr, name = [], None, None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            name = r[1]
            cur = []
            kernels.append((name, cur))
        elif r and r[0] == "Address":
            hdr = r
        elif cur is not None:
            cur.append(r)
    for name, data in kernels:
        if kernel_sub in name:
            return name, hdr, data
    raise SystemExit("kernel %r not in report (have %s)" % (kernel_sub, [k for k, _ in kernels]))


def line_table(kernel_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    lines = []
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
        in_fn, cur_line, cur_file = False, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                in_fn = re.search(kernel_re, m.group(1)) is not None
                continue
            if not in_fn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
                continue
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
                lines.append((cur_file, cur_line, ln.strip()))
    return lines


def main():
    rep, sub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    name, hdr, data = sass_rows(rep, sub)
    ix = {h: i for i, h in enumerate(hdr)}
    # mangled-name fragment for nvdisasm
    frag = sys.argv[4] if len(sys.argv) > 4 else re.sub(r"[^A-Za-z0-9_]", "", sub)
    table = line_table(frag)
    print("# %s: %d SASS rows in report, %d in cubin line table" % (name[:80], len(data), len(table)))
    n = min(len(data), len(table))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    by_line = defaultdict(lambda: defaultdict(int))
    tot = 0
    for i in range(n):
        r = data[i]
        key = (table[i][0], table[i][1])
        s = int(r[ix["# Samples"]])
        tot += s
        by_line[key]["samples"] += s
        by_line[key]["inst"] += int(r[ix["Instructions Executed"]])
        for c in stall_cols:
            by_line[key][c] += int(r[ix[c]])
    src_cache = {}

    def src(file, line):
        if file not in src_cache:
            path = os.path.join(ROOT, "optical-rl-gym_b200", "csrc", file or "")
            src_cache[file] = open(path).read().splitlines() if os.path.exists(path) else []
        ls = src_cache[file]
        return ls[line - 1].strip()[:90] if line and 0 < line <= len(ls) else ""

    print("# total samples %d" % tot)
    for key, d in sorted(by_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        stalls = sorted(((c[6:], d[c]) for c in stall_cols if d[c]), key=lambda x: -x[1])[:3]
        print("%5.1f%% %6d smp %9d inst  %s:%s  %-90s %s" % (100.0 * d["samples"] / max(tot, 1), d["samples"], d["inst"],
                                                            key[0], key[1], src(*key), stalls))


if __name__ == "__main__":
    main()
