#!/usr/bin/env python
"""bench.py — env agent-steps/sec of the batched DMFB env.step on B200 (BASELINE.json metric).

  python bench.py --gpus 1 --steps 20000 --warmup 200         # own arm, one JSON line on stdout
  python bench.py --impl reference --steps 3 --warmup 1       # CPU arm: the oracle port on all host threads
  torchrun --nproc-per-node N ... bench.py --gpus N ...       # one rank per GPU, envs sharded, no collective

Workload (config.workload): DMFB 10x10 chip, 4 droplets, fov 9, 65,536 envs per GPU, uniform random
actions (pre-generated, resident in HBM), auto-reset when an episode ends (all done or 40 steps), every
step's observation written to a rotating [41, N, A, 245] int8 episode buffer (2.6 GB > L2, so the stores
really reach HBM; no separate L2 flush).  One "step" = DMFBenv.step on all envs of the GPU = one launch of the
step kernel (auto-reset fused).

value     whole-job agent-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks
e2e       same metric through the HOST-buffer C ABI (dmfb_host_step): actions H2D from pinned memory,
          obs / reward / done / info D2H, every step, inside the timed region
roofline  step kernel alone: algorithmic bytes per launch / CUDA-event duration vs the measured HBM peak
cpu_baseline  the C oracle port of the reference env (oracle/) on this box's host threads
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, L, A, FOV = 10, 10, 4, 9
N_PER_GPU = 65536
D = 3 * FOV * FOV + 2
EP_LEN = 2 * (W + L)
METRIC = "env agent-steps/sec (DMFB 10x10 4d fov9, 64K envs/GPU)"
UNIT = "agent-steps/s"
# algorithmic bytes per env-step for C1 (obs 980 + reward 16 + team 4 + done 4 +
# avail 20 + info 5 + actions 4 + pos R/W 16 + goals 8 + counters R/W 12)
ALG_BYTES_PER_ENV_STEP = 1069


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20000)
    p.add_argument("--warmup", type=int, default=200)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--envs", type=int, default=N_PER_GPU, help="envs per GPU (default 65536)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--no-clocks", action="store_true")
    p.add_argument("--quick", action="store_true", help="short roofline / e2e legs (profiling runs under ncu)")
    p.add_argument("--no-others", action="store_true", help="skip the context timings of the other BASELINE configs")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--ref-seconds", type=float, default=20.0, help="--impl reference: wall time of the timed call")
    p.add_argument("--sub-batches", type=int, default=4,
                   help="the batch of every GPU is stepped as this many sub-batches on as many streams (1 = one launch per step)")
    p.add_argument("--window-ms", type=float, default=100.0, help="minimum length of one timed window")
    p.add_argument("--windows", type=int, default=7, help="timed windows; the median is reported")
    return p.parse_args()


# ------------------------------------------------------------------ helpers --
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(seconds, threads=None):
    """The C oracle port of DMFBenv.step (oracle/dmfb_oracle.c) on the host cores: bounded sample of the
    same workload (same chip, droplets, fov, random actions, auto-reset, full obs written every step)."""
    import oracle
    oracle.build()
    threads = threads or os.cpu_count() or 1
    n_envs = 256 * threads
    # calibrate: ~0.3 s probe, then size the real sample to `seconds`
    t0 = time.perf_counter()
    n, _ = oracle.dmfb_rollout(W, L, A, FOV, True, False, n_envs, 40, seed=1, threads=threads)
    dt = time.perf_counter() - t0
    steps = max(40, int(40 * seconds / max(dt, 1e-3)))
    steps = min(steps, 400000)
    t0 = time.perf_counter()
    n, _ = oracle.dmfb_rollout(W, L, A, FOV, True, False, n_envs, steps, seed=2, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_envs} envs x {steps} lock-step env-steps, C oracle port of the reference env "
                      f"(the reference itself is single-threaded Python, ~8e3 agent-steps/s/core; see BASELINE.md), "
                      f"{dt:.1f} s wall"}


# ------------------------------------------------------------ reference arm --
REF_PY_FILE = os.path.join(ROOT, "profiles", "reference_python_cpu.json")


def reference_python_context():
    """The reference's own Python env (unmodified, behind the import shim) timed in the BUILD container by
    tools/time_reference_python.py - it cannot travel to the GPU box (/root/reference does not exist there), so the
    committed measurement is quoted with its provenance, as context only."""
    try:
        with open(REF_PY_FILE) as f:
            return json.load(f)
    except Exception:
        return None


def run_reference(args, rank, world):
    """CPU arm: the C oracle port of DMFBenv.step on every host thread, same config as the own arm (65,536 chips,
    uniform random actions, auto-reset, the full observation written every env-step).  One bench "step" = S env-steps
    of all 65,536 chips (S sized so that the whole run takes ~20 s); the K timed steps run in ONE call, so that no
    thread start-up sits inside a step."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = os.cpu_count() or 1
    n_envs = args.envs
    t0 = time.perf_counter()
    oracle.dmfb_rollout(W, L, A, FOV, True, False, n_envs, 8, seed=3, threads=threads)       # calibration: the slope
    t1 = time.perf_counter()
    oracle.dmfb_rollout(W, L, A, FOV, True, False, n_envs, 48, seed=3, threads=threads)
    per_env_step = max((time.perf_counter() - t1) - (t1 - t0), 1e-4) / 40
    S = max(1, int(round(args.ref_seconds / max(args.steps, 1) / max(per_env_step, 1e-6))))
    S = min(S, 4000)
    if args.warmup > 0:
        oracle.dmfb_rollout(W, L, A, FOV, True, False, n_envs, min(S * args.warmup, 3 * S), seed=4, threads=threads)
    t0 = time.perf_counter()
    total, _ = oracle.dmfb_rollout(W, L, A, FOV, True, False, n_envs, S * args.steps, seed=10, threads=threads)
    dt = time.perf_counter() - t0
    value = total / dt
    sample = (f"each step = {S} env-steps of all {n_envs} chips ({S * n_envs * A} agent-steps), C oracle port of the "
              f"reference env on {threads} host threads, the {args.steps} steps in one call ({dt:.1f} s)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(n_envs, 1, extra={"sample": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "reference_python": reference_python_context()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_envs, world, extra=None):
    """`config` of both arms (kept identical so that the two lines describe the same workload; how each arm schedules
    the batch - host threads there, sub-batches on streams here - goes into `extra`)."""
    slots = EP_LEN
    cfg = {"workload": f"DMFB {W}x{L} chip, {A} droplets, fov {FOV}, {n_envs} envs/GPU, random actions, "
                       f"auto-reset with staggered episode phases, obs to rotating [{slots + 1},N,A,{D}] buffer",
           "envs_per_gpu": n_envs, "parallelism": f"env-shard x{world} (no collective)"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------ own arm --
def run_b200(args, rank, world, local_rank):
    import torch
    pkg = importlib.import_module("marl-dmfb_b200")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry exactly one
        # JSON line, so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = pkg._native.load()
    N = args.envs
    K = max(1, args.sub_batches)
    env = pkg.BatchedDMFB(N, W, L, A, fov=FOV, stall=True, b_degrade=False, device=dev, seed=1234,
                          env_base=rank * N, sub_batches=K)
    K = len(env._sub.ranges) if env._sub is not None else 1
    slots = EP_LEN
    obs_buf = torch.empty(slots + 1, N, A, D, dtype=torch.int8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    actions = torch.randint(0, 5, (slots, N, A), device=dev, generator=gen, dtype=torch.int8)
    env.reset(out=obs_buf[0])
    # steady state of a long rollout: episode phases spread uniformly (env n is n mod 40 steps into its episode),
    # so every step resets ~1/40 of the envs instead of all of them every 40th step
    env.step_count.copy_(torch.arange(N, device=dev, dtype=torch.int32) % EP_LEN)
    stream = torch.cuda.Stream(device=dev)

    def do_steps(t0, k):
        # consecutive steps are chained without a join: while one sub-batch waits for its own previous step to drain,
        # the kernels of the other sub-batches keep HBM busy (marl-dmfb_b200/pipeline.py); one join at the end
        for t in range(t0, t0 + k):
            s = t % slots
            env.step(actions[s], auto_reset=True, out=obs_buf[s + 1], join=False)
        env.join()

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- CUDA graphs: the step kernels are ~15 us, so a run is launch-bound unless they are replayed from a graph.
    # The timed region is a whole number R of passes of `--steps` steps: n_full replays of a graph of GRAPH_STEPS steps (a
    # multiple of the rotating buffer, so every replay continues where the last one ended) + one graph with the rest.
    GRAPH_STEPS = slots * 6
    graph_unit = graph_rem = None
    with torch.cuda.stream(stream):
        do_steps(0, max(args.warmup, 3))  # warm-up (also triggers cudaFuncSetAttribute / module load)
        stream.synchronize()
        if not args.no_graph:
            graph_unit = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_unit, stream=stream):
                do_steps(0, GRAPH_STEPS)
            graph_unit.replay()
            # calibration: how many steps fill one window
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(4):
                graph_unit.replay()
            c1.record(stream)
            stream.synchronize()
            est_ms_per_step = c0.elapsed_time(c1) / (4 * GRAPH_STEPS)
        else:
            est_ms_per_step = 0.03
    repeats = max(1, -(-int(args.window_ms / est_ms_per_step) // args.steps))      # ceil
    if dist is not None:     # same R on every rank
        r_all = torch.tensor([repeats], device=dev, dtype=torch.int64)
        dist.all_reduce(r_all, op=dist.ReduceOp.MAX)
        repeats = int(r_all.item())
    total_steps = args.steps * repeats
    n_full, rem = (total_steps // GRAPH_STEPS, total_steps % GRAPH_STEPS) if graph_unit is not None else (0, total_steps)
    if graph_unit is not None and rem:
        with torch.cuda.stream(stream):
            graph_rem = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_rem, stream=stream):
                do_steps(0, rem)
            graph_rem.replay()
            stream.synchronize()

    def timed_region():
        if graph_unit is None:
            do_steps(0, total_steps)
            return
        for _ in range(n_full):
            graph_unit.replay()
        if rem:
            graph_rem.replay()

    # Windows: the GPU is kept busy across ev0 (one untimed replay is queued first) and nothing synchronises with the
    # host between ev0 and ev1, so neither launch latency nor the host sits inside the window; `--windows` windows,
    # the median is reported.  nvidia-smi is polled on rank 0 only.
    barrier()
    launches0 = lib.dmfb_launch_count()
    clocks = ClockSampler(local_rank)
    sample_clocks = (not args.no_clocks) and rank == 0
    if sample_clocks:
        clocks.__enter__()
    window_ms = []
    with torch.cuda.stream(stream):
        for _ in range(max(1, args.windows)):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if graph_unit is not None:
                graph_unit.replay()                  # untimed: the window starts on a busy GPU
            ev0.record(stream)
            timed_region()
            ev1.record(stream)
            stream.synchronize()
            window_ms.append(ev0.elapsed_time(ev1))
    barrier()
    ms = statistics.median(window_ms)
    gpu_launches = total_steps * K  # one fused step(+auto-reset) kernel per sub-batch and step, per window
    _ = lib.dmfb_launch_count() - launches0
    t_all = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms_max = float(t_all.item())
    value = world * N * A * total_steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel: the step kernel alone (no auto-reset), CUDA events on its stream ----
    roof = None
    if rank == 0:
        peak, peak_src = measured_peak()
        with torch.cuda.stream(stream):
            for t in range(5):
                env.step(actions[t % slots], out=obs_buf[t % slots + 1])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=stream):
                for r in range(6):
                    for t in range(slots):
                        env.step(actions[t], out=obs_buf[t + 1], join=False)
                env.join()
            g2.replay()
            stream.synchronize()
            # ~0.6 s so that nvidia-smi samples clocks under load; short when the caller asked for a short run (ncu)
            reps = 2 if args.quick else int(0.6 / (6 * slots * 20e-6))
            e0.record(stream)
            for _ in range(reps):
                g2.replay()
            e1.record(stream)
            stream.synchronize()
        per_step_s = e0.elapsed_time(e1) * 1e-3 / (reps * 6 * slots)
        alg_bytes = ALG_BYTES_PER_ENV_STEP * N
        achieved = alg_bytes / per_step_s / 1e9
        # K launches (one per sub-batch) make one step of the batch; they overlap, so a launch has no duration of its own:
        # achieved = algorithmic bytes of ALL launches of the window / the window = bytes per launch / (window / launches)
        roof = {"bound": "hbm", "kernel": "dmfb_step_kernel<fov=9,G=4,A=4,E=16,deg=false>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_launch": alg_bytes // K, "us_per_launch": per_step_s * 1e6 / K,
                "launches_timed": reps * 6 * slots * K, "concurrent_launches": K,
                "alg_bytes_per_step": alg_bytes, "us_per_step": per_step_s * 1e6,
                "note": f"{K} launches (sub-batches of {N // K} envs on {K} streams) per step of the batch; us_per_launch = "
                        "window / launches.  A frac a little above 1 is possible: the peak is a COPY (read + write), this "
                        "kernel is 99 % writes"}
        tr = os.path.join(ROOT, "profiles", "traffic_step_kernel.json")
        if os.path.exists(tr):
            try:
                with open(tr) as f:
                    tj = json.load(f)
                    # measured on whole-batch launches (one launch = one step of 65,536 envs): per step of the batch
                    roof["traffic_per_step"] = tj.get("dram_bytes_per_launch")
                    roof["traffic"] = tj.get("dram_bytes_per_launch") / K
            except Exception:
                pass
    if sample_clocks:
        clocks.__exit__(None, None, None)
    env.reset()

    # ---- the other BASELINE configs, step kernel only (reported for context, not part of `value`) ----
    others = None
    if rank == 0 and not args.no_others:
        others = {}
        del obs_buf
        torch.cuda.empty_cache()

        def quick(make, n_act, alg_bytes, slots2=16):
            e2 = make()
            buf = torch.empty(slots2 + 1, e2.N, e2.A, e2.D, dtype=torch.int8, device=dev)
            acts = torch.randint(0, n_act, (slots2, e2.N, e2.A), device=dev, generator=gen, dtype=torch.int8)
            with torch.cuda.stream(stream):
                for t in range(3):
                    e2.step(acts[t], out=buf[t + 1])
                stream.synchronize()
                gq = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gq, stream=stream):
                    for t in range(slots2):
                        e2.step(acts[t], out=buf[t + 1], join=False)
                    e2.join()
                gq.replay()
                stream.synchronize()
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record(stream)
                for _ in range(10):
                    gq.replay()
                q1.record(stream)
                stream.synchronize()
            us = q0.elapsed_time(q1) * 1e3 / (10 * slots2)
            return {"us_per_step": us, "agent_steps_per_s": e2.N * e2.A / (us * 1e-6),
                    "alg_GBps": alg_bytes * e2.N / (us * 1e-6) / 1e9, "frac_of_peak": alg_bytes * e2.N / (us * 1e-6) / 1e9 / peak}

        others["C2 DMFB 20x20 10d fov9"] = quick(lambda: pkg.BatchedDMFB(N, 20, 20, 10, fov=9, device=dev, seed=1, sub_batches=K), 5, 2641, slots2=16)
        torch.cuda.empty_cache()
        others["C3 DMFB 50x50 10d fov9 degrade"] = quick(lambda: pkg.BatchedDMFB(N, 50, 50, 10, fov=9, b_degrade=True,
                                                                                  per_degrade=1.0, device=dev, seed=1,
                                                                                  sub_batches=K), 5, 2761, slots2=16)
        torch.cuda.empty_cache()
        # MEDA: without degradation nothing reads the usage counters, so they are not kept (BatchedMEDA default); the
        # "usage" / "degrade" lines add the counters (a 2-byte usage-log entry per droplet-step, replayed at reset)
        # and the 25-cell float64 health gather
        def meda(ver, **kw):
            return lambda: pkg.BatchedMEDA(N, 30, 60, 4, fov=19, obs_version=ver, device=dev, seed=1, sub_batches=K, **kw)

        others["C4 MEDA 30x60 4d fov19 (base obs, int8)"] = quick(meda(0), 9, 5897)
        torch.cuda.empty_cache()
        others["C4 MEDA 30x60 4d fov19 (v0_2 obs)"] = quick(meda(2), 9, 4453)
        torch.cuda.empty_cache()
        others["C4 MEDA 30x60 4d fov19 (base obs, usage counters kept)"] = quick(meda(0, track_usage=True), 9, 5897 + 8)
        torch.cuda.empty_cache()
        others["C4 MEDA 30x60 4d fov19 (base obs, degrade)"] = quick(meda(0, b_degrade=True, per_degrade=1.0), 9,
                                                                     5897 + 8 + 800)
        torch.cuda.empty_cache()
        obs_buf = None

    # ---- e2e: host-buffer C ABI, H2D actions + D2H results every step ----
    e2e = None
    if not args.no_e2e:
        import numpy as np
        obs_buf = None
        torch.cuda.empty_cache()
        # several ranks per host: each rank, its pinned buffers and its unpack threads on the GPU's own NUMA node
        numa_cpus = pkg.pin_to_gpu_numa(local_rank) if world > 1 else None
        henv = pkg.HostDMFB(N, W, L, A, fov=FOV, device=local_rank, seed=1234, env_base=rank * N, n_chunks=1)
        henv.reset()
        rng = np.random.default_rng(5 + rank)
        host_actions = [rng.integers(0, 5, (N, A)).astype(np.int8) for _ in range(4)]
        k_e2e = max(10, min(args.steps, 60))
        # The transfer of the 65.9 MB of observations is the whole cost of this path.  Pick, on this host, between the
        # plain DMA and the packed transfer (4-bit cells over PCIe, expanded by the host cores): a few untimed steps each.
        threads = len(numa_cpus) if numa_cpus else max(1, (os.cpu_count() or 1) // max(1, world))
        if numa_cpus:   # ranks that share a NUMA node share its cores
            import torch.distributed as _d
            same = [None] * world
            _d.all_gather_object(same, sorted(numa_cpus)[:1])
            threads = max(1, len(numa_cpus) // max(1, sum(1 for x in same if x == sorted(numa_cpus)[:1])))
        modes = [("plain DMA", 0, 100)]
        if threads >= 4 and A <= 15:
            modes += [(f"packed: 4-bit cells, {threads} host threads", threads, 0),
                      (f"packed: 4-bit cells, {threads} host threads, 20% of the envs by plain DMA", threads, 20)]
        best = None
        for label, th, pct in modes:
            henv.set_transfer(th, pct)
            for t in range(2):
                henv.step(host_actions[t % 4], auto_reset=True)
            t0 = time.perf_counter()
            for t in range(6):
                henv.step(host_actions[t % 4], auto_reset=True)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, label, th, pct)
        henv.set_transfer(best[2], best[3])
        for t in range(3):
            henv.step(host_actions[t % 4], auto_reset=True)
        barrier()
        t0 = time.perf_counter()
        for t in range(k_e2e):
            henv.step(host_actions[t % 4], auto_reset=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * N * A * k_e2e / float(t_e.item()), "unit": UNIT,
               "h2d_bytes_per_step": henv.h2d_bytes_per_step, "d2h_bytes_per_step": henv.d2h_bytes_per_step,
               "steps": k_e2e, "ms_per_step": float(t_e.item()) / k_e2e * 1e3,
               "api": "dmfb_host_step (pinned host buffers); transfer chosen on this host from " +
                      ", ".join(m[0] for m in modes) + ": " + best[1],
               "bound": "host memory: every step lands 64 MB of observations per GPU in host DRAM (PCIe Gen5 x16 carries "
                        "~55 GB/s; the packed transfer halves the PCIe bytes, the host cores then write the 64 MB)",
               "numa_pinned_cpus": len(numa_cpus) if numa_cpus else None}
        henv.close()

    if rank == 0:
        cpu = None if args.no_cpu else cpu_baseline(args.cpu_seconds)
        if cpu is not None:
            cpu["reference_python"] = reference_python_context()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_max / total_steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "repeats": repeats, "windows_ms": window_ms,
                "config": workload_config(N, world, extra={
                    "l2": f"outputs rotate over {(slots + 1) * N * A * D / 1e9:.2f} GB > L2 (no explicit flush)",
                    "sub_batches": f"every step of the {N} envs of a GPU is {K} launches (sub-batches of {N // K} envs on {K} "
                                   f"streams, chained without a join inside a graph): the kernels of one sub-batch fill the "
                                   f"bubble between two dependent launches of another",
                    "timing": f"{len(window_ms)} windows of {repeats} x {args.steps} steps = {total_steps} steps each "
                              f"(CUDA graphs of {GRAPH_STEPS} steps, GPU busy at the start event, no host sync inside), "
                              f"median window, max over ranks"}),
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": gpu_launches,
                "roofline": roof, "cpu_baseline": cpu, "other_configs": others}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
