#!/usr/bin/env python
"""CUDA-event timing of the DMFB step for an arbitrary configuration (experiment knob for DESIGN.md's breakdowns).
usage: python tools/time_variant.py W L A fov degrade(0|1) track_usage(0|1) [n_envs] [health<1 fraction] [obs_version]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("marl-dmfb_b200")
W, L, A, fov, deg, track = [int(x) for x in sys.argv[1:7]]
N = int(sys.argv[7]) if len(sys.argv) > 7 else 65536
frac = float(sys.argv[8]) if len(sys.argv) > 8 else 0.0
ver = int(sys.argv[9]) if len(sys.argv) > 9 else 0
env = pkg.BatchedDMFB(N, W, L, A, fov=fov, b_degrade=bool(deg), per_degrade=1.0, device="cuda:0", seed=1,
                      track_usage=bool(track), obs_version=ver)
env.reset(new=True)
if deg and frac > 0:
    g = torch.Generator(device="cuda:0").manual_seed(2)
    h = torch.rand(env.health.shape, device="cuda:0", generator=g, dtype=torch.float64)
    env.health.copy_(torch.where(h < frac, 0.5 + 0.5 * h, torch.ones_like(h)))
slots = 8
obs_buf = torch.empty(slots + 1, N, A, env.D, dtype=torch.int8, device="cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(1)
actions = torch.randint(0, 5, (slots, N, A), device="cuda:0", generator=gen, dtype=torch.int8)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(3):
        env.step(actions[i], out=obs_buf[i + 1])
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(slots):
            env.step(actions[i], out=obs_buf[i + 1])
    g.replay(); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(5):
        g.replay()
    e1.record(s)
    s.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * slots)
print(f"DMFB {W}x{L} A={A} fov={fov} deg={deg} usage={track} degraded-cells={frac} obs_version={ver}: step {us:8.2f} us")
