#!/usr/bin/env python
"""CUDA-event timing of the MEDA step at benchmark size (C4: 30x60 chip, 4 droplets, fov 19).
usage: python tools/time_meda.py [obs_version 0|1|2] [n_envs] [degrade 0|1] [usage 1|0] [auto_reset 0|1]
(usage 0: BatchedMEDA(track_usage=False), the default of a chip that does not degrade)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("marl-dmfb_b200")
ver = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
deg = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
W, L, A, fov = 30, 60, 4, 19
usage = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
env = pkg.BatchedMEDA(N, W, L, A, fov=fov, obs_version=ver, b_degrade=deg, per_degrade=1.0, device="cuda:0", seed=1,
                      track_usage=usage, sub_batches=int(os.environ.get("TK_SUB", "1")))
JOIN = int(os.environ.get("TK_SUB", "1")) == 1
auto = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
if auto:   # steady state of a long rollout: episode phases spread uniformly
    env.step_count.copy_(torch.arange(N, device="cuda:0", dtype=torch.int32) % env.max_step)
slots = 8
obs_buf = torch.empty(slots + 1, N, A, env.D, dtype=torch.int8, device="cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(1)
actions = torch.randint(0, 9, (slots, N, A), device="cuda:0", generator=gen, dtype=torch.int8)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(3):
        env.step(actions[i], out=obs_buf[i + 1], auto_reset=auto, join=JOIN)
    env.join()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(slots):
            env.step(actions[i], out=obs_buf[i + 1], auto_reset=auto, join=JOIN)
        env.join()
    g.replay(); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(5):
        g.replay()
    e1.record(s)
    s.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * slots)
bytes_env = A * env.D + 16 + 4 + 4 + 36 + 5 + 4 + 32 + 12
print(f"MEDA v{ver} deg={int(deg)} usage={int(usage or deg)} auto_reset={int(auto)} N={N}: step {us:8.2f} us  {N * A / us / 1e3:7.2f} G agent-steps/s  "
      f"{bytes_env * N / us / 1e3:8.1f} GB/s alg ({bytes_env} B/env-step)")
