#!/usr/bin/env python
"""Per-source-line instruction / stall-sample summary of an ncu report (needs -lineinfo and
--import-source on).  usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


cur, hdr, agg = None, None, collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 10 and cur and r[0].isdigit():
        d = dict(zip(hdr, r))
        k = (cur.split("/")[-1], int(r[0]), r[1][:110])
        a = agg.setdefault(k, [0, 0])
        a[0] += num(d["Instructions Executed"])
        a[1] += num(d["# Samples"])
tot = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print("total warp-instructions", tot, "samples", ts)
for (f, l, s), (n, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f:18s} {l:4d} {n:9d} {100 * n / tot:5.1f}%  smp {100 * sm / ts:5.1f}%  {s}")
