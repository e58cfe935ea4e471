#!/usr/bin/env python
"""Small driver for ncu: DMFB steps at the benchmark size exactly as bench.py issues them (fused auto-reset, staggered
episode phases, observations to a rotating buffer > L2), no CUDA graph, no timing.
usage: python tools/prof_step.py [c1|c2|c3] [steps] [noreset]"""
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
auto = not (len(sys.argv) > 3 and sys.argv[3] == "noreset")
c = CFG[name]
env = pkg.BatchedDMFB(c["N"], c["W"], c["L"], c["A"], fov=c["fov"], b_degrade=c["deg"], per_degrade=1.0,
                      device="cuda:0", seed=1234, sub_batches=int(os.environ.get("TK_SUB", "1")))
slots = max(8, min(steps, int(2.7e9 // (c["N"] * c["A"] * env.D))))
obs_buf = torch.empty(slots + 1, c["N"], c["A"], env.D, dtype=torch.int8, device="cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(1)
actions = torch.randint(0, 5, (slots, c["N"], c["A"]), device="cuda:0", generator=gen, dtype=torch.int8)
env.reset(out=obs_buf[0])
env.step_count.copy_(torch.arange(c["N"], device="cuda:0", dtype=torch.int32) % env.max_step)
for t in range(steps):
    env.step(actions[t % slots], auto_reset=auto, out=obs_buf[t % slots + 1], join=False)
env.join()
torch.cuda.synchronize()
print("ok", name, steps, "auto_reset" if auto else "no reset", int(env.step_count.max()))
