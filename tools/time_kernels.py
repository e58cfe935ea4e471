#!/usr/bin/env python
"""CUDA-event timing of the individual entry points at benchmark size.
usage: python tools/time_kernels.py [c1|c2|c3] [n_envs] [short]   (short: the two step timings only)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("marl-dmfb_b200")
CFG = {"c1": dict(W=10, L=10, A=4, fov=9, deg=False), "c2": dict(W=20, L=20, A=10, fov=9, deg=False),
       "c3": dict(W=50, L=50, A=10, fov=9, deg=True)}
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
c = CFG[name]
T = 2 * (c["W"] + c["L"])
env = pkg.BatchedDMFB(N, c["W"], c["L"], c["A"], fov=c["fov"], b_degrade=c["deg"], per_degrade=1.0,
                      device="cuda:0", seed=1234, usage_log=os.environ.get("TK_USAGE_LOG", "1") != "0",
                      task_prefetch=os.environ.get("TK_PREFETCH", "1") != "0",
                      health_bitmap=os.environ.get("TK_BITMAP", "1") != "0",
                      sub_batches=int(os.environ.get("TK_SUB", "1")))
JOIN = int(os.environ.get("TK_SUB", "1")) == 1
slots = min(T, max(8, int(2.6e9 // (N * c["A"] * env.D))))
obs_buf = torch.empty(slots + 1, N, c["A"], env.D, dtype=torch.int8, device="cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(1)
actions = torch.randint(0, 5, (slots, N, c["A"]), device="cuda:0", generator=gen, dtype=torch.int8)
alg = {"c1": 1069, "c2": 2641, "c3": 2761}[name]


def timeit(fn, reps):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def graph_time(fn, n):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(3):
            fn(i)
        env.join()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n):
                fn(i)
            env.join()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * n)


env.reset()
us = graph_time(lambda i: env.step(actions[i % slots], out=obs_buf[i % slots + 1], join=JOIN), slots)
print(f"{name} N={N}: step (no reset, graph)       {us:8.2f} us  {alg * N / us / 1e3:8.1f} GB/s alg")
if not (len(sys.argv) > 3 and sys.argv[3] == "short"):
    env.reset()
    us = graph_time(lambda i: env.step(actions[i % slots], auto_reset=True, out=obs_buf[i % slots + 1], join=JOIN), T)
    print(f"{name} N={N}: step (auto-reset, graph of T) {us:8.2f} us  {N * c['A'] / us / 1e3:8.2f} G agent-steps/s")
env.reset()
env.step_count.copy_(torch.arange(N, device="cuda:0", dtype=torch.int32) % T)   # stagger the episode phases
us = graph_time(lambda i: env.step(actions[i % slots], auto_reset=True, out=obs_buf[i % slots + 1], join=JOIN), T)
print(f"{name} N={N}: step (auto-reset, staggered)  {us:8.2f} us  {N * c['A'] / us / 1e3:8.2f} G agent-steps/s")
if len(sys.argv) > 3 and sys.argv[3] == "short":
    sys.exit(0)
us = timeit(lambda i: env.reset(out=obs_buf[i % slots]), 20)
print(f"{name} N={N}: reset all (device generator)  {us:8.2f} us")
lay = env.drop.clone()
us = timeit(lambda i: env.reset(layouts=lay, out=obs_buf[i % slots]), 20)
print(f"{name} N={N}: reset all (injected layouts)  {us:8.2f} us")
us = timeit(lambda i: env.get_obs(out=obs_buf[i % slots]), 20)
print(f"{name} N={N}: observe                       {us:8.2f} us  {N * c['A'] * env.D / us / 1e3:8.1f} GB/s")
st = torch.empty(4, N, 3, c["W"], c["L"], dtype=torch.int8, device="cuda:0") if N * 3 * c["W"] * c["L"] * 4 < 8e9 else None
if st is not None:
    us = timeit(lambda i: env.get_state(out=st[i % 4]), 20)
    print(f"{name} N={N}: get_state                     {us:8.2f} us  {N * 3 * c['W'] * c['L'] / us / 1e3:8.1f} GB/s")
