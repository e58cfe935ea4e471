"""HostDMFB — the DMFB hot path behind HOST buffers (dmfb_host_* of include/dmfb_b200.h).

Same call shape as the reference env (actions in host memory in, observations / rewards / dones /
info in host memory out, env/DMFB/dmfb.py:560-587), batched over N chips.  The handle owns the device
state; each call moves its inputs H2D and its results D2H.  This is the path bench.py reports as `e2e`."""
import ctypes as C

import numpy as np

from . import _native as nat


def gpu_cpu_affinity(device, topo_text=None):
    """CPUs of the NUMA node GPU `device` hangs off (`nvidia-smi topo -m`, column "CPU Affinity"), or None.
    `topo_text`: the command's output, if the caller already has it."""
    out = topo_text
    if out is None:
        import subprocess
        try:
            out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=10).stdout
        except Exception:
            return None
    import re
    out = re.sub(r"\x1b\[[0-9;]*m", "", out)                 # the header row is underlined with ANSI codes
    lines = [ln for ln in out.splitlines() if ln.strip()]
    if not lines or "CPU Affinity" not in lines[0]:
        return None
    header = [h.strip() for h in lines[0].split("\t") if h.strip()]
    try:
        col = header.index("CPU Affinity") + 1          # data rows start with the GPU name
    except ValueError:
        return None
    for ln in lines[1:]:
        f = [x.strip() for x in ln.split("\t")]
        f = [x for x in f if x != ""]
        if f and f[0] == f"GPU{device}" and len(f) > col:
            cpus = set()
            for part in f[col].split(","):
                try:
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
                except ValueError:
                    return None
            return cpus or None
    return None


def pin_to_gpu_numa(device):
    """Restrict this process to the CPUs next to GPU `device`, so that the pinned host buffers it allocates from now on
    (first touch) and the unpack threads live on that NUMA node.  Several ranks per host otherwise share one node's
    memory controllers for all their D2H traffic.  Returns the CPU set used, or None when the topology is unknown or
    the allowed set would become empty."""
    import os
    cpus = gpu_cpu_affinity(device)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except (AttributeError, OSError):
        return None


class _Pinned:
    """numpy array over cudaHostAlloc'ed memory (freed with the object)."""

    def __init__(self, lib, shape, dtype):
        self._lib = lib
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = lib.dmfb_host_alloc_pinned(max(nbytes, 16))
        if not self._ptr:
            raise MemoryError("cudaHostAlloc failed")
        buf = (C.c_uint8 * max(nbytes, 16)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            self._lib.dmfb_host_free_pinned(self._ptr)
        except Exception:
            pass


class HostDMFB:
    def __init__(self, n_envs, width, length, n_agents, n_blocks=0, fov=5, stall=True, b_degrade=False,
                 per_degrade=0.1, device=0, seed=0, env_base=0, n_chunks=1):
        self.lib = nat.load()
        self.cfg = nat.DmfbCfg()
        nat.check(self.lib.dmfb_cfg_init(C.byref(self.cfg), width, length, n_agents, n_blocks, fov, int(bool(stall)),
                                         int(bool(b_degrade)), float(per_degrade)), "dmfb_cfg_init")
        self.cfg.env_base = int(env_base)
        self.N, self.A, self.D = int(n_envs), n_agents, self.cfg.obs_dim
        self.seed = int(seed)
        self._h = C.c_void_p()
        nat.check(self.lib.dmfb_host_create(C.byref(self.cfg), self.N, int(device), int(n_chunks), C.byref(self._h)),
                  "dmfb_host_create")
        N, A = self.N, self.A
        self._pins = {k: _Pinned(self.lib, s, d) for k, (s, d) in dict(
            actions=((N, A), np.int8), obs=((N, A, self.D), np.int8), reward=((N, A), np.float32),
            done=((N, A), np.uint8), constraints=((N,), np.int32), success=((N,), np.uint8)).items()}
        self.actions = self._pins["actions"].array   # write your actions here (pinned) or pass an array to step()
        self.obs = self._pins["obs"].array
        self.reward = self._pins["reward"].array
        self.done = self._pins["done"].array
        self.constraints = self._pins["constraints"].array
        self.success = self._pins["success"].array
        self.h2d_bytes_per_step = N * A
        self.d2h_bytes_per_step = N * A * self.D + N * A * 4 + N * A + N * 4 + N

    def _p(self, a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def set_transfer(self, n_threads, dma_percent=50):
        """Packed observation transfer (dmfb_host_set_transfer): `dma_percent` % of the envs arrive unpacked by DMA,
        the rest as 4-bit cells expanded by `n_threads` host threads meanwhile.  n_threads=0: plain DMA."""
        nat.check(self.lib.dmfb_host_set_transfer(self._h, int(n_threads), int(dma_percent)), "dmfb_host_set_transfer")
        N, A, D = self.N, self.A, self.D
        small = N * A * 4 + N * A + N * 4 + N
        if n_threads <= 0:
            self.d2h_bytes_per_step = N * A * D + small
        else:   # mirrors host_step_packed(): DMA share in 64-env units, the rest as packed records
            n_dma = N if dma_percent == 100 else (N * int(dma_percent) // 100) & ~63
            stride = (((D - 2) + 1) // 2 + 2 + 3) & ~3
            self.d2h_bytes_per_step = n_dma * A * D + (N - n_dma) * A * stride + small

    def reset(self, new=False, layouts=None, degrade=None):
        lay = None if layouts is None else np.ascontiguousarray(layouts, np.uint8)
        deg = None if degrade is None else np.ascontiguousarray(degrade, np.float64)
        nat.check(self.lib.dmfb_host_reset(self._h, int(bool(new)), self._p(lay), self._p(deg), self.seed,
                                           self._p(self.obs)), "dmfb_host_reset")
        return self.obs

    def step(self, actions=None, draws=None, record=True, auto_reset=False):
        if actions is not None and actions is not self.actions:
            self.actions[...] = actions
        u = None if draws is None else np.ascontiguousarray(draws, np.float64)
        flags = (nat.STEP_RECORD_USAGE if record else 0) | (nat.STEP_AUTO_RESET if auto_reset else 0)
        nat.check(self.lib.dmfb_host_step(self._h, self._p(self.actions), self._p(u), self.seed, flags,
                                          self._p(self.obs), self._p(self.reward), self._p(self.done),
                                          self._p(self.constraints), self._p(self.success)), "dmfb_host_step")
        return self.obs, self.reward, self.done.view(np.bool_), {"constraints": self.constraints, "success": self.success}

    def close(self):
        if self._h:
            self.lib.dmfb_host_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostMEDA:
    """The MEDA hot path behind HOST buffers (meda_host_* of include/dmfb_b200.h): MEDAEnv.step / reset (meda.py:513-550)
    for N chips with numpy arrays in and out.  `constraints` is the punish count of the step (the reference's
    info['constraints'] is -0.6 * count)."""

    n_actions = 9

    def __init__(self, n_envs, width, length, n_agents, fov=19, b_degrade=False, per_degrade=0.1, obs_version=2,
                 device=0, seed=0, env_base=0):
        self.lib = nat.load()
        self.cfg = nat.MedaCfg()
        rc = self.lib.meda_cfg_init(C.byref(self.cfg), width, length, n_agents, fov, int(bool(b_degrade)),
                                    float(per_degrade), int(obs_version))
        if rc == 2:  # meda.py:151-154
            raise RuntimeError("Too many droplets in the " + str(width) + "x" + str(length) + " MEDA array")
        nat.check(rc, "meda_cfg_init")
        self.cfg.env_base = int(env_base)
        self.N, self.A, self.D = int(n_envs), n_agents, self.cfg.obs_dim
        self.max_step = self.cfg.max_step
        self.seed = int(seed)
        self._h = C.c_void_p()
        nat.check(self.lib.meda_host_create(C.byref(self.cfg), self.N, int(device), C.byref(self._h)), "meda_host_create")
        N, A = self.N, self.A
        self._pins = {k: _Pinned(self.lib, s, d) for k, (s, d) in dict(
            actions=((N, A), np.int8), obs=((N, A, self.D), np.int8), reward=((N, A), np.float32),
            done=((N, A), np.uint8), constraints=((N,), np.int32), success=((N,), np.uint8)).items()}
        for k, v in self._pins.items():
            setattr(self, k, v.array)
        self.h2d_bytes_per_step = N * A
        self.d2h_bytes_per_step = N * A * self.D + N * A * 4 + N * A + N * 4 + N

    def _p(self, a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def reset(self, new_chip=False, layouts=None, degrade=None):
        lay = None if layouts is None else np.ascontiguousarray(layouts, np.uint8)
        deg = None if degrade is None else np.ascontiguousarray(degrade, np.float64)
        nat.check(self.lib.meda_host_reset(self._h, int(bool(new_chip)), self._p(lay), self._p(deg), self.seed,
                                           self._p(self.obs)), "meda_host_reset")
        return self.obs

    def step(self, actions=None, draws=None, auto_reset=False):
        if actions is not None and actions is not self.actions:
            self.actions[...] = actions
        u = None if draws is None else np.ascontiguousarray(draws, np.float64)
        flags = nat.STEP_AUTO_RESET if auto_reset else 0
        nat.check(self.lib.meda_host_step(self._h, self._p(self.actions), self._p(u), self.seed, flags,
                                          self._p(self.obs), self._p(self.reward), self._p(self.done),
                                          self._p(self.constraints), self._p(self.success)), "meda_host_step")
        return self.obs, self.reward, self.done.view(np.bool_), {"constraints": self.constraints, "success": self.success}

    def close(self):
        if self._h:
            self.lib.meda_host_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
