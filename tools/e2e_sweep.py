import importlib, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
P = importlib.import_module("marl-dmfb_b200")
N, A = 65536, 4
rng = np.random.default_rng(0)
acts = [rng.integers(0, 5, (N, A)).astype(np.int8) for _ in range(4)]
for nc in (1, 2, 4, 8, 16):
    h = P.HostDMFB(N, 10, 10, 4, fov=9, device=0, seed=1, n_chunks=nc)
    h.reset()
    for t in range(3): h.step(acts[t % 4], auto_reset=True)
    t0 = time.perf_counter()
    for t in range(40): h.step(acts[t % 4], auto_reset=True)
    dt = (time.perf_counter() - t0) / 40
    print(nc, "chunks:", round(dt * 1e3, 3), "ms/step", round(h.d2h_bytes_per_step / dt / 1e9, 1), "GB/s", round(N * A / dt / 1e6, 1), "M agent-steps/s", flush=True)
    h.close()
