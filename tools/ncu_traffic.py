#!/usr/bin/env python
"""dram bytes per launch from an ncu report -> profiles/traffic_step_kernel.json (read by bench.py)."""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
def val(r, k):
    v = float(r[h.index(k)]); u = units[h.index(k)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
per = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in rows[2:]]
dur = [float(r[h.index("gpu__time_duration.sum")]) for r in rows[2:]]
json.dump({"dram_bytes_per_launch": sum(per) / len(per), "launches": len(per), "per_launch": per,
           "gpu_time_us": dur, "source": rep.split("/")[-1],
           "note": "ncu --set full --cache-control none --clock-control none, steady state (L2 already full of earlier steps' output)"},
          open(out, "w"), indent=1)
print(open(out).read())
