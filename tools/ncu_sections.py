#!/usr/bin/env python
"""Executed warp-instructions per source-line range of an ncu report.
usage: python tools/ncu_sections.py report.ncu-rep file.cu:lo-hi[:label] ..."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
def num(x):
    try: return int(float(x))
    except ValueError: return 0
cur = hdr = None
per = collections.Counter(); smp = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 10 and cur and r[0].isdigit():
        d = dict(zip(hdr, r)); per[(cur, int(r[0]))] += num(d["Instructions Executed"]); smp[(cur, int(r[0]))] += num(d["# Samples"])
tot = sum(per.values()) or 1; ts = sum(smp.values()) or 1
print("total", tot)
for spec in sys.argv[2:]:
    parts = spec.split(":"); f = parts[0]; lo, hi = map(int, parts[1].split("-")); label = parts[2] if len(parts) > 2 else spec
    n = sum(v for (ff, l), v in per.items() if ff == f and lo <= l <= hi); s = sum(v for (ff, l), v in smp.items() if ff == f and lo <= l <= hi)
    print(f"{label:28s} {n:9d} {100*n/tot:5.1f}%   samples {100*s/ts:5.1f}%")
other = [(k, v) for k, v in per.items() if k[0] != "dmfb_kernels.cu"]
print("other files:", sum(v for k, v in other), collections.Counter({k[0]: 0 for k, v in other}).keys())
