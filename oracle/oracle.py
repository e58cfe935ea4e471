"""ctypes binding of the CPU oracle (oracle/*.c) — TEST INFRASTRUCTURE ONLY.

The arrays mirror the product's struct-of-arrays boundary (include/dmfb_b200.h) so that parity
tests can compare buffers directly, but the usage matrix is float64 here, as in the reference
(dmfb.py:148), not uint16.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """gcc the oracle into oracle/liboracle.so (idempotent)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_dmfb_rollout.restype = C.c_int64
        if hasattr(_LIB, "orc_meda_rollout"):
            _LIB.orc_meda_rollout.restype = C.c_int64
    return _LIB


class _DmfbCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("width", "length", "n_agents", "fov", "stall", "b_degrade", "n_blocks",
                                         "obs_version")]


def _p(a, ct=None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class OracleDMFB:
    """N independent DMFB chips stepped by the C restatement of env/DMFB/dmfb.py.
    obs_version 0 = DMFBenv.getOneObs, 1 = DMFBenv_v0_1.getOneObs (int8 layers, direction entries = integer
    numerators of the reference's floats)."""

    def __init__(self, n_envs, width, length, n_agents, fov=9, stall=True, b_degrade=False, n_blocks=0, obs_version=0):
        self.N, self.W, self.L, self.A, self.fov = n_envs, width, length, n_agents, fov
        self.D = (4 if obs_version == 1 else 3) * fov * fov + 2
        self.n_blocks = n_blocks
        self.obs_version = obs_version
        self.cfg = _DmfbCfg(width, length, n_agents, fov, int(stall), int(b_degrade), n_blocks, obs_version)
        self.blocks = np.zeros((n_envs, n_blocks, 2), np.uint8) if n_blocks else None
        self.b_degrade = bool(b_degrade)
        N, A = n_envs, n_agents
        self.drop = np.zeros((N, A, 4), np.uint8)
        self.step_count = np.zeros(N, np.int32)
        self.constraints = np.zeros(N, np.int32)
        self.usage = np.zeros((N, width, length), np.float64)
        self.health = np.ones((N, width, length), np.float64)
        self.degrade = np.ones((N, width, length), np.float64)

    def with_version(self, obs_version):
        """observation of the current state under the other obs variant (state is shared, not copied)"""
        o = OracleDMFB.__new__(OracleDMFB)
        o.__dict__.update(self.__dict__)
        o.obs_version = obs_version
        o.D = (4 if obs_version == 1 else 3) * self.fov * self.fov + 2
        o.cfg = _DmfbCfg(self.W, self.L, self.A, self.fov, self.cfg.stall, self.cfg.b_degrade, self.n_blocks, obs_version)
        return o

    def reset(self, layouts, new=False, degrade=None, mask=None, blocks=None):
        obs = np.zeros((self.N, self.A, self.D), np.int8)
        layouts = np.ascontiguousarray(layouts, np.uint8)
        if blocks is not None:
            blocks = np.ascontiguousarray(blocks, np.uint8)
        if degrade is not None:
            degrade = np.ascontiguousarray(degrade, np.float64)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        lib().orc_dmfb_reset(C.byref(self.cfg), self.N, _p(mask), int(new), _p(layouts), _p(blocks), _p(degrade),
                             _p(self.drop), _p(self.blocks), _p(self.step_count), _p(self.constraints), _p(self.usage),
                             _p(self.health), _p(self.degrade), _p(obs))
        return obs

    def step(self, actions, draws=None, record=True, want_obs=True):
        N, A = self.N, self.A
        actions = np.ascontiguousarray(actions, np.int8)
        if draws is not None:
            draws = np.ascontiguousarray(draws, np.float64)
        obs = np.zeros((N, A, self.D), np.int8) if want_obs else None
        reward = np.zeros((N, A), np.float64)
        done = np.zeros((N, A), np.uint8)
        cons = np.zeros(N, np.int32)
        succ = np.zeros(N, np.uint8)
        rc = lib().orc_dmfb_step(C.byref(self.cfg), N, _p(self.drop), _p(self.blocks), _p(self.step_count), _p(self.constraints),
                                 _p(self.usage), _p(self.health) if self.b_degrade else None, _p(actions),
                                 _p(draws), int(record), _p(obs), _p(reward), _p(done), _p(cons), _p(succ))
        if rc:
            raise TypeError("action is illegal")
        return obs, reward, done, cons, succ

    def observe(self):
        obs = np.zeros((self.N, self.A, self.D), np.int8)
        lib().orc_dmfb_observe(C.byref(self.cfg), self.N, _p(self.drop), _p(self.blocks), _p(obs))
        return obs

    def global_state(self):
        out = np.zeros((self.N, 3, self.W, self.L), np.int8)
        lib().orc_dmfb_global_state(C.byref(self.cfg), self.N, _p(self.drop), _p(self.blocks), _p(out))
        return out

    def gen_layouts(self, seed):
        """Reference-distributed random tasks (dmfb.py:207-226) for all N chips."""
        out = np.zeros((self.N, self.A, 4), np.uint8)
        for e in range(self.N):
            s = C.c_uint64((seed << 20) + e)
            lib().orc_dmfb_gen_layout(C.byref(self.cfg), C.byref(s), _p(out[e]))
        return out

    def gen_blocks(self, seed, layouts):
        """Reference-distributed random 2x2 blocks (dmfb.py:228-251) for the given tasks."""
        layouts = np.ascontiguousarray(layouts, np.uint8)
        out = np.zeros((self.N, self.n_blocks, 2), np.uint8)
        for e in range(self.N):
            s = C.c_uint64((seed << 21) + e)
            lib().orc_dmfb_gen_blocks(C.byref(self.cfg), C.byref(s), _p(layouts[e]), _p(out[e]))
        return out


def dmfb_rollout(width, length, n_agents, fov, stall, b_degrade, n_envs, steps, seed=1, threads=1):
    """Timed CPU leg: returns agent-steps executed (see orc_dmfb_rollout)."""
    cfg = _DmfbCfg(width, length, n_agents, fov, int(stall), int(b_degrade), 0, 0)
    obs = np.zeros((n_envs, n_agents, 3 * fov * fov + 2), np.int8)
    chk = C.c_uint64(0)
    n = lib().orc_dmfb_rollout(C.byref(cfg), n_envs, steps, C.c_uint64(seed), _p(obs), int(threads), C.byref(chk))
    return int(n), int(chk.value)


class _MedaCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("width", "length", "n_agents", "fov", "b_degrade", "obs_version")]


class OracleMEDA:
    """N independent MEDA chips stepped by the C restatement of env/MEDA/meda.py.
    obs_version 0 = MEDAEnv.getOneObs (int8-cast), 1 = MEDAEnv_v0_1.getOneObs (int8 layers, direction
    entries = integer numerators of the reference's floats), 2 = MEDAEnv_v0_2.getOneObs."""

    def __init__(self, n_envs, width, length, n_agents, fov=19, b_degrade=False, obs_version=0):
        self.N, self.W, self.L, self.A, self.fov = n_envs, width, length, n_agents, fov
        self.obs_version = obs_version
        self.D = (3 if obs_version == 2 else 4) * fov * fov + 2
        self.cfg = _MedaCfg(width, length, n_agents, fov, int(b_degrade), obs_version)
        self.b_degrade = bool(b_degrade)
        N, A = n_envs, n_agents
        self.drop = np.zeros((N, A, 4), np.uint8)
        self.status = np.zeros((N, A), np.uint8)
        self.step_count = np.zeros(N, np.int32)
        self.fails = np.zeros(N, np.int32)
        self.usage = np.zeros((N, width, length), np.float64)
        self.health = np.ones((N, width, length), np.float64)
        self.degrade = np.ones((N, width, length), np.float64)

    def with_version(self, obs_version):
        """observation of the current state under the other obs variant (state is shared, not copied)"""
        o = OracleMEDA.__new__(OracleMEDA)
        o.__dict__.update(self.__dict__)
        o.obs_version = obs_version
        o.D = (3 if obs_version == 2 else 4) * self.fov * self.fov + 2
        o.cfg = _MedaCfg(self.W, self.L, self.A, self.fov, int(self.b_degrade), obs_version)
        return o

    def reset(self, layouts, mask=None):
        obs = np.zeros((self.N, self.A, self.D), np.int8)
        layouts = np.ascontiguousarray(layouts, np.uint8)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        lib().orc_meda_reset(C.byref(self.cfg), self.N, _p(mask), _p(layouts), _p(self.drop), _p(self.status),
                             _p(self.step_count), _p(self.fails), _p(self.usage), _p(self.health), _p(self.degrade),
                             _p(obs))
        return obs

    def step(self, actions, draws=None, want_obs=True):
        N, A = self.N, self.A
        actions = np.ascontiguousarray(actions, np.int8)
        if draws is not None:
            draws = np.ascontiguousarray(draws, np.float64)
        obs = np.zeros((N, A, self.D), np.int8) if want_obs else None
        reward = np.zeros((N, A), np.float64)
        done = np.zeros((N, A), np.uint8)
        cons = np.zeros(N, np.int32)
        succ = np.zeros(N, np.uint8)
        lib().orc_meda_step(C.byref(self.cfg), N, _p(self.drop), _p(self.status), _p(self.step_count), _p(self.fails),
                            _p(self.usage), _p(self.health) if self.b_degrade else None, _p(actions), _p(draws),
                            _p(obs), _p(reward), _p(done), _p(cons), _p(succ))
        return obs, reward, done, cons, succ

    def observe(self):
        obs = np.zeros((self.N, self.A, self.D), np.int8)
        lib().orc_meda_observe(C.byref(self.cfg), self.N, _p(self.drop), _p(obs))
        return obs

    def gen_layouts(self, seed):
        out = np.zeros((self.N, self.A, 4), np.uint8)
        for e in range(self.N):
            s = C.c_uint64((seed << 20) + e)
            lib().orc_meda_gen_layout(C.byref(self.cfg), C.byref(s), _p(out[e]))
        return out


def cpython_set_order(mask_bits, n_max):
    order = (C.c_int * 64)()
    n = lib().orc_cpython_set_order(C.c_uint32(mask_bits), int(n_max), order)
    return list(order[:n])


def meda_rollout(width, length, n_agents, fov, b_degrade, obs_version, n_envs, steps, seed=1, threads=1):
    cfg = _MedaCfg(width, length, n_agents, fov, int(b_degrade), obs_version)
    D = (3 if obs_version == 2 else 4) * fov * fov + 2
    obs = np.zeros((n_envs, n_agents, D), np.int8)
    chk = C.c_uint64(0)
    n = lib().orc_meda_rollout(C.byref(cfg), n_envs, steps, C.c_uint64(seed), _p(obs), int(threads), C.byref(chk))
    return int(n), int(chk.value)
