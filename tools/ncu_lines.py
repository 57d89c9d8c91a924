#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line:
   python tools/ncu_lines.py report.ncu-rep <kernel regex> [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = defaultdict(lambda: [0, 0, ""])
stall = defaultdict(lambda: defaultdict(int))
tot = 0
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        i_x = hdr.index("Instructions Executed")
        st_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= i_x or r[0] == "":  # SASS rows repeat the samples of their source line
        continue
    try:
        s = int(r[i_s]); x = int(r[i_x])
    except ValueError:
        continue
    key = r[0]
    agg[key][0] += s; agg[key][1] += x; agg[key][2] = r[1]
    tot += s
    for k, h in st_cols:
        try:
            stall[key][h] += int(r[k])
        except ValueError:
            pass
print("total samples", tot)
for key, (s, x, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    ss = sorted(stall[key].items(), key=lambda kv: -kv[1])[:3]
    print("%5s %6.2f%% inst=%9d  %-90s %s" % (key, 100.0 * s / max(tot, 1), x, src.strip()[:90], " ".join("%s:%d" % (h[6:], v) for h, v in ss if v)))
