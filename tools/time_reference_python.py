#!/usr/bin/env python
"""Times the UNMODIFIED reference env (env/DMFB/dmfb.py, env/MEDA/meda.py of /root/reference) on the host cores of
the BUILD container - BASELINE.md section 3: (i) one process, (ii) P = os.cpu_count() independent processes with one
env each (the reference has no vectorisation or multiprocessing of its own, so that is its best case).

The reference cannot travel to the GPU box, so the result is committed as profiles/reference_python_cpu.json and
bench.py quotes it, tagged with this provenance, as `cpu_baseline.reference_python` (context, not a bench number).

    python tools/time_reference_python.py [seconds per measurement, default 10]
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

CONFIGS = {
    "C1 DMFB 10x10 4d fov9": ("dmfb", dict(width=10, length=10, n_agents=4, fov=9)),
    "C2 DMFB 20x20 10d fov9": ("dmfb", dict(width=20, length=20, n_agents=10, fov=9)),
    "C3 DMFB 50x50 10d fov9 degrade": ("dmfb", dict(width=50, length=50, n_agents=10, fov=9, b_degrade=True, per_degrade=1.0)),
    "C4 MEDA 30x60 4d fov19 (base obs)": ("meda", dict(w=30, l=60, n_agents=4, fov=19)),
    "C4 MEDA 30x60 4d fov19 (v0_2 obs)": ("meda2", dict(w=30, l=60, n_agents=4, fov=19)),
}


def worker(args):
    kind, kw, seconds, seed = args
    import numpy as np
    import ref_shim
    ref_dmfb, ref_meda = ref_shim.install()
    np.random.seed(seed)
    if kind == "dmfb":
        env = ref_dmfb.DMFBenv(kw["width"], kw["length"], kw["n_agents"], 0, **{k: v for k, v in kw.items()
                                                                                  if k not in ("width", "length", "n_agents")})
        n_act = 5
    else:
        cls = ref_meda.MEDAEnv if kind == "meda" else ref_meda.MEDAEnv_v0_2
        env = cls(kw["w"], kw["l"], kw["n_agents"], fov=kw["fov"])
        n_act = 9
    A = kw["n_agents"]
    rng = np.random.default_rng(seed)
    env.reset()
    steps, t_reset, t_step = 0, 0.0, 0.0
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        acts = [int(a) for a in rng.integers(0, n_act, A)]
        t0 = time.perf_counter()
        obs, rew, done, info = env.step(acts)
        t_step += time.perf_counter() - t0
        steps += 1
        if all(done.values()):
            t0 = time.perf_counter()
            env.reset()
            t_reset += time.perf_counter() - t0
    return steps * A, t_step, t_reset


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    P = os.cpu_count() or 1
    cpu = platform.processor() or ""
    try:
        with open("/proc/cpuinfo") as f:
            cpu = [ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")][0]
    except Exception:
        pass
    out = {"provenance": "unmodified reference env from /root/reference behind tests/golden/ref_shim.py, timed in the build "
                         "container (NOT on the GPU box: the Python reference cannot travel) by tools/time_reference_python.py",
           "cpu": cpu, "cores_visible": P, "python": platform.python_version(), "seconds_per_measurement": seconds,
           "unit": "agent-steps/s", "note": "uniform random actions, auto-reset at episode end; step time only "
                                            "(reset time reported separately as ms per reset-call)",
           "configs": {}}
    ctx = mp.get_context("fork")
    for name, (kind, kw) in CONFIGS.items():
        n1, ts1, tr1 = worker((kind, kw, seconds, 1))
        with ctx.Pool(P) as pool:
            t0 = time.perf_counter()
            res = pool.map(worker, [(kind, kw, seconds, 10 + k) for k in range(P)])
            wall = time.perf_counter() - t0
        out["configs"][name] = {
            "one_process": n1 / ts1, "one_process_incl_resets": n1 / (ts1 + tr1),
            "P_processes": sum(r[0] / r[1] for r in res), "P": P,
            "P_processes_wall_incl_resets_and_startup": sum(r[0] for r in res) / wall}
        print(name, json.dumps(out["configs"][name]), flush=True)
    path = os.path.join(ROOT, "profiles", "reference_python_cpu.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
