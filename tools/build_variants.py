#!/usr/bin/env python
"""Builds experimental variants of the library next to the product build: lib/variants/<name>.so, one per set of -D
flags, selected at run time with DMFB_B200_LIB=<path> (see _native.lib_path).  Kernel experiments that need a GPU
round trip can then be compared in ONE gpurun call.
usage: python tools/build_variants.py name=FLAG1,FLAG2 [name2=...]   (flags without the -D)"""
import importlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("marl-dmfb_b200.build")
out_dir = os.path.join(b.PKG, "lib", "variants")
os.makedirs(out_dir, exist_ok=True)


def one(spec):
    name, _, flags = spec.partition("=")
    out = os.path.join(out_dir, name + ".so")
    defs = ["-D" + f for f in flags.split(",") if f]
    cmd = ["nvcc"] + b.NVCC_FLAGS + defs + ["-I", os.path.join(ROOT, "include"), "-o", out] + b.sources()
    subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
    return out


with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)
