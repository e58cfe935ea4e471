#!/usr/bin/env python
"""Small driver for ncu: a few DMFB steps at the benchmark size, no CUDA graph, no timing.
usage: python tools/prof_step.py [config] [steps]   config in c1 (default), c2, c3"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("marl-dmfb_b200")

CFG = {"c1": dict(W=10, L=10, A=4, fov=9, deg=False, N=65536),
       "c2": dict(W=20, L=20, A=10, fov=9, deg=False, N=65536),
       "c3": dict(W=50, L=50, A=10, fov=9, deg=True, N=65536)}
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
c = CFG[name]
env = pkg.BatchedDMFB(c["N"], c["W"], c["L"], c["A"], fov=c["fov"], b_degrade=c["deg"], per_degrade=1.0,
                      device="cuda:0", seed=1234)
obs_buf = torch.empty(steps + 1, c["N"], c["A"], env.D, dtype=torch.int8, device="cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(1)
actions = torch.randint(0, 5, (steps, c["N"], c["A"]), device="cuda:0", generator=gen, dtype=torch.int8)
env.reset(out=obs_buf[0])
for t in range(steps):
    env.step(actions[t], auto_reset=True, out=obs_buf[t + 1])
torch.cuda.synchronize()
print("ok", name, steps, int(env.step_count.max()))
