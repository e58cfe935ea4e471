#!/usr/bin/env python
"""Small pass over every kernel family for compute-sanitizer (memcheck): a few steps / resets of small batches.
usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = importlib.import_module("marl-dmfb_b200")
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)


def run_dmfb(N, W, L, A, fov, nb=0, deg=False, ver=0, K=1, steps=None):
    env = P.BatchedDMFB(N, W, L, A, nb, fov=fov, b_degrade=deg, per_degrade=1.0, device=dev, seed=3, obs_version=ver,
                        sub_batches=K, reward_f64=True)
    if deg:
        env.usage.fill_(49)
    T = steps or (2 * (W + L) + 3)
    for t in range(T):
        acts = torch.randint(0, 5, (N, A), device=dev, generator=g, dtype=torch.int8)
        env.step(acts, auto_reset=True, join=(t % 3 == 0))
    env.join()
    env.step(acts, freeze_terminated=True)
    env.reset(mask=(torch.arange(N, device=dev) % 2).to(torch.uint8))
    env.restart()
    env.get_obs(); env.get_state(); env.usage_counts() if deg else None
    env.check()
    torch.cuda.synchronize()
    print("dmfb ok", N, W, L, A, fov, nb, deg, ver, K, flush=True)


def run_meda(N, W, L, A, ver, deg, K=1):
    env = P.BatchedMEDA(N, W, L, A, fov=19, obs_version=ver, b_degrade=deg, per_degrade=1.0, device=dev, seed=4, sub_batches=K)
    if deg:
        env.usage.fill_(48)
    for t in range(W + L + 3):
        acts = torch.randint(0, 9, (N, A), device=dev, generator=g, dtype=torch.int8)
        env.step(acts, auto_reset=True, join=(t % 2 == 0))
    env.join()
    env.reset(mask=(torch.arange(N, device=dev) % 2).to(torch.uint8))
    env.restart(); env.get_obs(); env.check()
    if deg:
        env.usage_counts()
    torch.cuda.synchronize()
    print("meda ok", N, W, L, A, ver, deg, K, flush=True)


run_dmfb(100, 10, 10, 4, 9)                       # C1 instance, partial last tile
run_dmfb(333, 10, 10, 4, 9, deg=True, K=3)        # C1 + degradation, sub-batches
run_dmfb(77, 20, 20, 10, 9)                       # C2 instance + task-search kernel
run_dmfb(50, 50, 50, 10, 9, deg=True, K=2, steps=210)   # C3 instance
run_dmfb(61, 12, 15, 6, 7, nb=3, deg=True)        # generic instance with obstacles
run_dmfb(40, 16, 11, 5, 8)                        # even fov
run_dmfb(30, 14, 14, 4, 7, ver=1)                 # v0_1 observation
run_dmfb(9, 30, 30, 20, 11)                       # 32-lane groups
run_meda(65, 30, 60, 4, 0, False)
run_meda(65, 30, 60, 4, 2, True, K=2)
run_meda(33, 45, 60, 6, 1, True)
run_meda(17, 80, 80, 10, 2, False)
h = P.HostDMFB(200, 10, 10, 4, fov=9, device=0, seed=1, n_chunks=2)
h.reset()
for t in range(45):
    h.step(np.random.default_rng(t).integers(0, 5, (200, 4)).astype(np.int8), auto_reset=True)
h.set_transfer(2, 50)
for t in range(5):
    h.step(np.random.default_rng(t).integers(0, 5, (200, 4)).astype(np.int8), auto_reset=True)
h.close()
m = P.HostMEDA(50, 30, 60, 4, fov=19, obs_version=2, b_degrade=True, per_degrade=1.0, device=0, seed=2)
m.reset(new_chip=True)
for t in range(95):
    m.step(np.random.default_rng(t).integers(0, 9, (50, 4)).astype(np.int8), auto_reset=True)
m.close()
print("sanitize smoke: all ok")
