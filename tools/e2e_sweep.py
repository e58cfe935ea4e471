#!/usr/bin/env python
"""Sweep of the host-buffer path (dmfb_host_step) at the bench size: chunk streams, and the packed transfer at
several DMA / packed splits and host thread counts.  usage: python tools/e2e_sweep.py [n_envs]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = importlib.import_module("marl-dmfb_b200")
N, A = (int(sys.argv[1]) if len(sys.argv) > 1 else 65536), 4
rng = np.random.default_rng(0)
acts = [rng.integers(0, 5, (N, A)).astype(np.int8) for _ in range(4)]


def run(label, n_chunks=1, threads=0, pct=100):
    h = P.HostDMFB(N, 10, 10, A, fov=9, device=0, seed=1, n_chunks=n_chunks)
    if threads:
        h.set_transfer(threads, pct)
    h.reset()
    for t in range(3):
        h.step(acts[t % 4], auto_reset=True)
    t0 = time.perf_counter()
    for t in range(40):
        h.step(acts[t % 4], auto_reset=True)
    dt = (time.perf_counter() - t0) / 40
    print(f"{label:34s} {dt * 1e3:7.3f} ms/step {h.d2h_bytes_per_step / dt / 1e9:6.1f} GB/s into host buffers "
          f"{N * A / dt / 1e6:7.1f} M agent-steps/s", flush=True)
    h.close()


for nc in (1, 2, 8):
    run(f"plain DMA, {nc} chunk stream(s)", n_chunks=nc)
cores = os.cpu_count() or 1
for th in sorted({cores, max(1, cores // 2)}, reverse=True):
    for pct in (0, 20, 30, 40, 50, 60):
        run(f"packed, {th} threads, {pct}% by DMA", threads=th, pct=pct)
