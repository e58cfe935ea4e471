#!/usr/bin/env python
"""Experiment: the batch stepped as K independent sub-batches on K streams (one CUDA graph with K parallel chains), so
that the bubble between two dependent launches of one sub-batch is filled by the other sub-batches' kernels.
usage: python tools/time_streams.py [c1|c2] [n_envs] [K ...]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("marl-dmfb_b200")
CFG = {"c1": dict(W=10, L=10, A=4, fov=9, alg=1069), "c2": dict(W=20, L=20, A=10, fov=9, alg=2641)}
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
Ks = [int(x) for x in sys.argv[3:]] or [1, 2, 4]
c = CFG[name]
T = 2 * (c["W"] + c["L"])
dev = torch.device("cuda:0")
for auto in (False, True):
    for K in Ks:
        n = N // K
        envs = [pkg.BatchedDMFB(n, c["W"], c["L"], c["A"], fov=c["fov"], device=dev, seed=1234, env_base=k * n) for k in range(K)]
        D = envs[0].D
        slots = T
        obs = torch.empty(slots + 1, N, c["A"], D, dtype=torch.int8, device=dev)
        gen = torch.Generator(device=dev).manual_seed(1)
        actions = torch.randint(0, 5, (slots, N, c["A"]), device=dev, generator=gen, dtype=torch.int8)
        for k, e in enumerate(envs):
            e.reset(out=obs[0, k * n:(k + 1) * n])
            if auto:
                e.step_count.copy_(torch.arange(k * n, (k + 1) * n, device=dev, dtype=torch.int32) % T)
        main = torch.cuda.Stream()
        subs = [torch.cuda.Stream() for _ in range(K)]

        def chain(k, steps):
            lo, hi = k * n, (k + 1) * n
            for t in range(steps):
                envs[k].step(actions[t % slots, lo:hi], auto_reset=auto, out=obs[t % slots + 1, lo:hi])

        with torch.cuda.stream(main):
            for k in range(K):
                chain(k, 3)
            main.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main):
                fork = torch.cuda.Event()
                fork.record(main)
                for k in range(K):
                    subs[k].wait_event(fork)
                    with torch.cuda.stream(subs[k]):
                        chain(k, T)
                        j = torch.cuda.Event()
                        j.record(subs[k])
                    main.wait_event(j)
            g.replay(); main.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for _ in range(8):
                g.replay()
            e1.record(main)
            main.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (8 * T)
        print(f"{name} N={N} auto_reset={int(auto)} K={K} sub-batches: {us:7.2f} us per step of the whole batch  "
              f"{c['alg'] * N / us / 1e3:7.1f} GB/s alg  {N * c['A'] / us / 1e3:6.2f} G agent-steps/s", flush=True)
        del envs, obs, actions, g
        torch.cuda.empty_cache()
