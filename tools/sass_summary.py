#!/usr/bin/env python
"""Per-kernel SASS summary of the built library (no GPU needed): instruction count, TMA bulk stores (UBLKCP), atomics,
shared / local memory accesses, VABSDIFF4, programmatic-dependent-launch markers, registers and stack.
usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "marl-dmfb_b200", "lib", "libdmfb_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731


def short(name):
    d = demangle(name)
    d = re.sub(r"\(anonymous namespace\)::|dmfb::|void |<unnamed>::|\((int|bool)\)", "", d)
    m = re.match(r"([A-Za-z_0-9]+(<[^>]*>)?)", d)
    return m.group(1) if m else d


regs = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+)", ln)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
stats = collections.OrderedDict()
cur = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        c = stats[cur]
        c["instr"] += 1
        for key, pat in (("UBLKCP", r"^UBLKCP"), ("RED/ATOM", r"^(RED|ATOM|ATOMG|ATOMS)"), ("STS", r"^STS"), ("LDS", r"^LDS"),
                         ("LDL+STL", r"^(LDL|STL)"), ("VABSDIFF4", r"^VABSDIFF4"), ("PDL", r"^(ACQBULK|PREEXIT)"),
                         ("SHFL", r"^SHFL"), ("tensor", r"^(HMMA|UTC|LDTM|STTM)")):
            if re.match(pat, op):
                c[key] += 1
print(f"# SASS summary of marl-dmfb_b200/lib/libdmfb_b200.so (cuobjdump -sass / -res-usage, sm_100a)")
print("# UBLKCP = TMA bulk store of a finished observation tile (cp.async.bulk.global.shared::cta); PDL = ACQBULK / PREEXIT")
print("# (programmatic dependent launch); tensor = HMMA / UTC*MMA / LDTM / STTM (none: nothing here is a contraction).")
print("# instr | UBLKCP | RED/ATOM | STS | LDS | SHFL | VABSDIFF4 | LDL+STL | PDL | tensor | regs | stack | kernel")
rows = []
for k, c in stats.items():
    r = regs.get(k, (0, 0))
    rows.append((short(k), c, r))
for name, c, r in sorted(rows):
    print(f"{c['instr']:6d} | {c['UBLKCP']:2d} | {c['RED/ATOM']:3d} | {c['STS']:4d} | {c['LDS']:3d} | {c['SHFL']:3d} | {c['VABSDIFF4']:3d} | "
          f"{c['LDL+STL']:3d} | {c['PDL']:2d} | {c['tensor']:1d} | {r[0]:3d} | {r[1]:4d} | {name}")
