"""Host-side mirror of the reference MEDA env interface (env/MEDA/meda.py) over GPU-resident state.

* ``BatchedMEDA``  — N independent MEDA chips in HBM, stepped by the sm_100a kernels through the C ABI.
* ``MEDAEnv`` / ``MEDAEnv_v0_2`` — N = 1 adapters with the reference's Python types.  ``MEDAEnv`` returns the
  base observation (float64, length 4*fov^2+2, meda.py:613-674), ``MEDAEnv_v0_2`` the int8 one (3*fov^2+2,
  meda.py:850-897).

Deliberate deviation (SURVEY.md note A): ``get_env_info()['obs_shape']`` is the DMFB-style tuple
``(C, fov, fov, 2, C*fov*fov+2)`` — the reference's base class returns an int there, which crashes every consumer
(vdn.py:12, replay_buffer.py:10).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .pipeline import SubBatches, offset_ptr


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def build_set_order_table(n_agents):
    """[2^A, A] uint8: row m = iteration order of the python set {i: bit i of m} built by ascending insertion
    (what `for idx in observed` sees, meda.py:862-872), 0xFF padded.  Built with REAL python sets and cross-checked
    against the library's own CPython-set emulation (meda_set_order)."""
    lib = nat.load()
    tab = np.full((1 << n_agents, n_agents), 0xFF, np.uint8)
    row = (C.c_uint8 * n_agents)()
    for m in range(1 << n_agents):
        s = set()
        for i in range(n_agents):
            if (m >> i) & 1:
                s.add(i)
        order = list(s)
        tab[m, :len(order)] = order
        nat.check(lib.meda_set_order(m, n_agents, row), "meda_set_order")
        if list(row) != list(tab[m]):
            raise RuntimeError(f"CPython set order differs from the library's emulation for mask {m:#x}: "
                               f"{order} vs {list(row)}")
    return tab


class BatchedMEDA:
    """N chips x A droplets (5x5 micro-electrode squares).  reset/step/get_obs/get_avail_actions/get_env_info."""

    n_actions = 9

    def __init__(self, n_envs, width, length, n_agents, fov=19, b_degrade=False, per_degrade=0.1, obs_version=2,
                 device="cuda", seed=0, env_base=0, reward_f64=False, degrade=None, layouts=None, track_usage=None,
                 usage_log=True, reset_list=False, health_bitmap=True, sub_batches=1):
        # track_usage: keep the m_usage actuation counters (addUsage, meda.py:591-598).  Nothing reads them unless the
        # chip degrades (updateHealth runs only `if self.b_degrade`, meda.py:547-548), so like BatchedDMFB the default
        # is `b_degrade`; the N=1 adapters always track them because `m_usage` is a visible attribute there.
        self.lib = nat.load()
        self.cfg = nat.MedaCfg()
        rc = self.lib.meda_cfg_init(C.byref(self.cfg), width, length, n_agents, fov, int(bool(b_degrade)),
                                    float(per_degrade), int(obs_version))
        if rc == 2:  # meda.py:151-154
            raise RuntimeError("Too many droplets in the " + str(width) + "x" + str(length) + " MEDA array")
        nat.check(rc, "meda_cfg_init")
        self.cfg.env_base = int(env_base)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedMEDA needs a CUDA device: there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.N, self.W, self.L, self.A, self.fov = int(n_envs), width, length, n_agents, fov
        self.obs_version = int(obs_version)
        self.D = self.cfg.obs_dim
        self.max_step = self.cfg.max_step
        self.b_degrade = bool(b_degrade)
        self.seed = int(seed)
        self.agents = ["player_{}".format(i) for i in range(n_agents)]
        N, A, dev = self.N, self.A, self.device
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        self.drop = z(N, A, 4, dtype=torch.uint8)          # x_center, y_center, goal x_center, goal y_center
        self.start = z(N, A, 2, dtype=torch.uint8)
        self.status = z(N, A, dtype=torch.uint8)
        self.step_count = z(N, dtype=torch.int32)
        self.fails = z(N, dtype=torch.int32)               # punish count; the reference's `fails` is -0.6 * this
        self.terminated = z(N, dtype=torch.uint8)
        self.episode = z(N, dtype=torch.int32)
        if track_usage is None:
            track_usage = self.b_degrade
        self.usage = z(N, width, length, dtype=torch.int32) if (track_usage or self.b_degrade) else None
        self._health = torch.ones(N, width, length, dtype=torch.float64, device=dev) if self.b_degrade else None
        # degraded-cell bit map (meda_state_t.health_bits): 25 clear bits under a droplet skip the float64 gather.  Kept
        # in step with `health` by the kernels; reading `env.health` marks it stale (the tensor may be written through)
        # and the next call rebuilds it first.
        self._health_bits = (z(N, (width * length + 31) // 32, dtype=torch.int32)
                             if (self.b_degrade and health_bitmap) else None)
        self._health_dirty = False
        self.degrade = torch.ones(N, width, length, dtype=torch.float64, device=dev) if self.b_degrade else None
        # steps log the actuated droplets; resets and usage_counts() fold the log into `usage` (meda_state_t.usage_log)
        self._usage_log = bool(usage_log) and self.usage is not None
        self.usage_log = z(N, self.max_step, A, dtype=torch.int16) if self._usage_log else None
        self.usage_log_len = z(N, dtype=torch.int32) if self._usage_log else None
        # auto_reset is fused into the step kernel (the warp that stepped an env resets it); reset_list=True: the step
        # only lists the envs that terminated and a second, small kernel resets exactly those (
        # a masked reset sweeps the whole batch after every step instead)
        self.reset_list = z(N, dtype=torch.int32) if reset_list else None
        self.reset_count = z(2, dtype=torch.int32) if reset_list else None
        self.gen_status = z(1, dtype=torch.int32)      # sticky DMFB_STATUS_SAMPLER_GAVE_UP
        self.state = self._make_state(0, N)
        self.set_order = None
        if self.obs_version != nat.MEDA_OBS_BASE and A > 16:
            # the others' goals are painted in CPython-set iteration order (meda.py:862-878), served from a [2^A][A] table
            raise ValueError("MEDAEnv_v0_1 / _v0_2 observations are supported for at most 16 droplets")
        if self.obs_version != nat.MEDA_OBS_BASE and A > 8:
            self.set_order = torch.as_tensor(build_set_order_table(A)).to(dev)
        self.obs = z(N, A, self.D, dtype=torch.int8)
        self.reward = z(N, A, dtype=torch.float32)
        self.reward_f64 = z(N, A, dtype=torch.float64) if reward_f64 else None
        self.team_reward = z(N, dtype=torch.float32)
        self.done = z(N, A, dtype=torch.uint8)
        self.avail = torch.ones(N, A, self.n_actions, dtype=torch.uint8, device=dev)
        self.constraints = z(N, dtype=torch.int32)
        self.success = z(N, dtype=torch.uint8)
        self.term_out = z(N, dtype=torch.uint8)
        self.padded = z(N, dtype=torch.uint8)
        self._out = self._make_out(self.obs)
        # sub_batches: K > 1 steps the batch as K sub-batches on K streams (pipeline.py)
        self._sub = SubBatches(self.device, N, sub_batches) if int(sub_batches) > 1 and not reset_list else None
        if self._sub is not None:
            self._sub_cfg, self._sub_state = [], []
            for lo, hi in self._sub.ranges:
                cfg = nat.MedaCfg.from_buffer_copy(self.cfg)
                cfg.env_base = self.cfg.env_base + lo
                self._sub_cfg.append(cfg)
                self._sub_state.append(self._make_state(lo, hi))
            self._sub_out = [self._make_out(self.obs, lo) for lo, _ in self._sub.ranges]
        self.reset(new_chip=True, layouts=layouts, degrade=degrade)

    def _make_state(self, lo, hi):
        """meda_state_t over the envs [lo, hi) of the batch (pointers offset into the same tensors)."""
        p = offset_ptr
        whole = lo == 0 and hi == self.N
        return nat.MedaState(
            n_envs=hi - lo, usage_log_cap=self.max_step if self._usage_log else 0,
            drop=p(self.drop, lo), start=p(self.start, lo), status=p(self.status, lo), step_count=p(self.step_count, lo),
            fails=p(self.fails, lo), terminated=p(self.terminated, lo), episode=p(self.episode, lo),
            usage=p(self.usage, lo), health=p(self._health, lo), degrade=p(self.degrade, lo),
            usage_log=p(self.usage_log, lo), usage_log_len=p(self.usage_log_len, lo),
            reset_list=self.reset_list.data_ptr() if (whole and self.reset_list is not None) else None,
            reset_count=self.reset_count.data_ptr() if (whole and self.reset_count is not None) else None,
            gen_status=self.gen_status.data_ptr(), health_bits=p(self._health_bits, lo))

    def _make_out(self, obs, lo=0):
        p = offset_ptr
        return nat.MedaOut(
            obs=p(obs, lo), reward=p(self.reward, lo), reward_f64=p(self.reward_f64, lo),
            team_reward=p(self.team_reward, lo), done=p(self.done, lo), avail=p(self.avail, lo),
            constraints=p(self.constraints, lo), success=p(self.success, lo),
            terminated=p(self.term_out, lo), padded=p(self.padded, lo), status=None)

    def join(self):
        """Makes the caller's stream wait for sub-batch steps issued with join=False (no-op otherwise)."""
        if self._sub is not None:
            self._sub.join()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def health(self):
        """m_health [N,W,L] float64 (None without degradation); may be written through."""
        if self._health_bits is not None:
            self._health_dirty = True
        return self._health

    def _sync_health(self):
        if self._health_dirty:
            self.join()
            self._health_dirty = False
            with torch.cuda.device(self.device):
                rc = self.lib.meda_sync_health_bits(C.byref(self.cfg), C.byref(self.state), self._stream())
            nat.check(rc, "meda_sync_health_bits")

    def _as(self, t, dtype, shape, name):
        if t is None:
            return None
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.ascontiguousarray(t))
        t = t.to(device=self.device, dtype=dtype).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def reset(self, mask=None, layouts=None, new_chip=False, degrade=None, out=None):
        """MEDAEnv.reset (meda.py:541-550): new tasks, observation, then updateHealth.
        layouts: optional [N,A,4] (x_c, y_c, goal x_c, goal y_c) with centres in [2, dim-3]."""
        self.join()
        mask_t = self._as(mask, torch.uint8, (self.N,), "mask")
        lay_t = self._as(layouts, torch.uint8, (self.N, self.A, 4), "layouts")
        deg_t = self._as(degrade, torch.float64, (self.N, self.W, self.L), "degrade")
        obs = self.obs if out is None else out
        self._sync_health()
        with torch.cuda.device(self.device):
            rc = self.lib.meda_reset(C.byref(self.cfg), C.byref(self.state), _ptr(mask_t), int(bool(new_chip)),
                                     _ptr(lay_t), _ptr(deg_t), self.seed, _ptr(self.set_order), _ptr(obs),
                                     self._stream())
        nat.check(rc, "meda_reset")
        return obs

    def restart(self, mask=None):
        """MEDAEnv.restart (meda.py:552-561): droplets back to their start squares; `fails` is kept."""
        self.join()
        mask_t = self._as(mask, torch.uint8, (self.N,), "mask")
        with torch.cuda.device(self.device):
            rc = self.lib.meda_restart(C.byref(self.cfg), C.byref(self.state), _ptr(mask_t), _ptr(self.set_order),
                                       _ptr(self.obs), self._stream())
        nat.check(rc, "meda_restart")
        return self.obs

    def step(self, actions, draws=None, freeze_terminated=False, auto_reset=False, out=None, join=True):
        """MEDAEnv.step (meda.py:513-539) on every env; actions [N,A] in 0..8."""
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.ascontiguousarray(actions))
        if actions.device != self.device:
            actions = actions.to(self.device)
        if actions.dtype not in (torch.int8, torch.uint8, torch.int32, torch.int64):
            actions = actions.to(torch.int64)
        actions = actions.contiguous()
        if tuple(actions.shape) != (self.N, self.A):
            raise RuntimeError("The number of actions is not the same as n_droplets")  # meda.py:242-244
        draws_t = self._as(draws, torch.float64, (self.N, self.A), "draws")
        flags = (nat.STEP_FREEZE_TERM if freeze_terminated else 0) | (nat.STEP_AUTO_RESET if auto_reset else 0)
        obs, o = (self.obs, self._out) if out is None else (out, self._make_out(out))
        self._sync_health()
        if self._sub is None:
            with torch.cuda.device(self.device):
                rc = self.lib.meda_step(C.byref(self.cfg), C.byref(self.state), _ptr(actions), actions.element_size(),
                                        _ptr(draws_t), self.seed, flags, _ptr(self.set_order), C.byref(o), self._stream())
            nat.check(rc, "meda_step")
        else:       # K sub-batches on K streams; join=False lets the next step's kernels overlap these (pipeline.py)
            sub, es = self._sub, actions.element_size()
            sub.fork(keep_alive=(actions, draws_t, out))
            with torch.cuda.device(self.device):
                for k, ((lo, hi), stream) in enumerate(zip(sub.ranges, sub.handles())):
                    o_k = self._sub_out[k] if out is None else self._make_out(out, lo)
                    rc = self.lib.meda_step(C.byref(self._sub_cfg[k]), C.byref(self._sub_state[k]),
                                            C.c_void_p(actions.data_ptr() + lo * self.A * es), es,
                                            None if draws_t is None else C.c_void_p(draws_t.data_ptr() + lo * self.A * 8),
                                            self.seed, flags, _ptr(self.set_order), C.byref(o_k), stream)
                    nat.check(rc, "meda_step")
            if join:
                sub.join()
        info = {"constraints": self.constraints, "success": self.success, "terminated": self.term_out.view(torch.bool),
                "team_reward": self.team_reward, "padded": self.padded.view(torch.bool)}
        return obs, self.reward, self.done.view(torch.bool), info

    def get_obs(self, out=None):
        self.join()
        obs = self.obs if out is None else out
        with torch.cuda.device(self.device):
            rc = self.lib.meda_observe(C.byref(self.cfg), C.byref(self.state), _ptr(self.set_order), _ptr(obs),
                                       self._stream())
        nat.check(rc, "meda_observe")
        return obs

    def get_avail_actions(self):
        return self.avail

    def check(self):
        """RuntimeError if the task generator gave up since the last check (droplets that cannot be placed; the
        reference would loop for ever, meda.py:213-233; the env kept its previous layout).  Device -> host sync."""
        self.join()
        if int(self.gen_status.item()) & nat.STATUS_SAMPLER_GAVE_UP:
            self.gen_status.zero_()
            raise RuntimeError("the task generator found no legal droplet placement for this chip")

    def get_env_info(self):
        c = 3 if self.obs_version == nat.MEDA_OBS_V02 else 4
        return {"n_actions": self.n_actions, "n_agents": self.A, "obs_shape": (c, self.fov, self.fov, 2, self.D),
                "episode_limit": self.max_step}

    def usage_counts(self):
        """m_usage (folds the usage log into the counters first)."""
        self.join()
        if self._usage_log:
            with torch.cuda.device(self.device):
                rc = self.lib.meda_flush_usage(C.byref(self.cfg), C.byref(self.state), self._stream())
            nat.check(rc, "meda_flush_usage")
        return self.usage


class _MedaRoutingView:
    def __init__(self, env):
        self._e = env

    @property
    def status(self):
        return [bool(x) for x in self._e._b.status[0].cpu().numpy()]

    @property
    def centers(self):
        return self._e._b.drop[0, :, 0:2].cpu().numpy().astype(int)

    @property
    def goal_centers(self):
        return self._e._b.drop[0, :, 2:4].cpu().numpy().astype(int)


class MEDAEnv:
    """Drop-in for env.MEDA.meda.MEDAEnv (meda.py:457-681) on top of a 1-env GPU batch."""

    metadata = {"render.modes": ["human", "rgb_array"]}
    _obs_version = nat.MEDA_OBS_BASE

    def __init__(self, w, l, n_agents, n_blocks=0, fov=19, stall=True, b_degrade=False, per_degrade=0.1, show=False,
                 savemp4=False, device="cuda", seed=None, layouts=None, degrade=None):
        assert w > 0 and l > 0
        assert n_agents > 0
        if seed is None:
            seed = int(np.random.randint(0, 2**31 - 1))
        self._b = BatchedMEDA(1, w, l, n_agents, fov=fov, b_degrade=b_degrade, per_degrade=per_degrade,
                              obs_version=self._obs_version, device=device, seed=seed, reward_f64=True, track_usage=True,
                              layouts=None if layouts is None else np.asarray(layouts)[None],
                              degrade=None if degrade is None else np.asarray(degrade)[None])
        self.agents = list(self._b.agents)
        self.possible_agents = self.agents[:]
        self.width, self.length, self.fov = w, l, fov
        self.b_degrade = b_degrade
        self.max_step = self._b.max_step
        self.mode = None
        self.rewards = {a: 0.0 for a in self.agents}
        self.dones = {a: False for a in self.agents}
        self._fails_f = 0
        self.routing_manager = _MedaRoutingView(self)

    # reference attributes
    @property
    def step_count(self):
        return int(self._b.step_count[0].item())

    @property
    def fails(self):
        """The reference's running float `self.fails += fail` (meda.py:521), accumulated in its order."""
        return self._fails_f

    def _punish_sum(self):
        """`np.sum(calPunish())` (meda.py:256,321-330) in the reference's accumulation order: -0.6 subtracted once per
        close pair from both droplets, then summed over the droplets - so that `info['constraints']` and `fails` are
        the reference's floats to the last bit, not just -0.6 * count."""
        d = self._b.drop[0, :, 0:2].cpu().numpy().astype(np.int64)
        n = len(self.agents)
        punish = [0] * n
        for i in range(n - 1):
            for j in range(i + 1, n):
                if int(((d[i] - d[j]) ** 2).sum()) < 36:      # centre distance < 1.5 * (r_i + r_j) = 6
                    punish[i] -= 0.6
                    punish[j] -= 0.6
        return np.sum(punish)

    @property
    def m_health(self):
        b = self._b
        return np.ones((b.W, b.L)) if b.health is None else b.health[0].cpu().numpy()

    @property
    def m_usage(self):
        return self._b.usage_counts()[0].cpu().numpy().astype(np.float64)

    @property
    def m_degrade(self):
        b = self._b
        return np.ones((b.W, b.L)) if b.degrade is None else b.degrade[0].cpu().numpy()

    def _obs_list(self, obs):
        o = obs[0].cpu().numpy()
        if self._obs_version != nat.MEDA_OBS_V02:
            o = o.astype(np.float64)   # the base class and v0_1 return float64 (meda.py:621,673 / :795,841)
        if self._obs_version == nat.MEDA_OBS_V01:
            # the kernel emits the numerators; dir_VEC = (dy / width, dx / length) (meda.py:840)
            o[:, -2] = o[:, -2] / self.width
            o[:, -1] = o[:, -1] / self.length
        return [o[i].copy() for i in range(len(self.agents))]

    def step(self, actions):
        if isinstance(actions, dict):
            acts = [actions[a] for a in self.agents]
        elif isinstance(actions, list):
            acts = actions
        else:
            raise UnboundLocalError("acts")                      # meda.py:516-520 leaves `acts` unbound
        if len(acts) != len(self.agents):
            raise RuntimeError("The number of actions is not the same as n_droplets")
        a = torch.as_tensor(np.asarray([int(x) for x in acts], dtype=np.int64)[None])
        obs, _, done, info = self._b.step(a)
        r = self._b.reward_f64[0].cpu().numpy()
        d = done[0].cpu().numpy()
        for k, name in enumerate(self.agents):
            self.rewards[name] = float(r[k])
            self.dones[name] = bool(d[k])
        fail = self._punish_sum()
        assert round(float(fail) / -0.6) == int(info["constraints"][0].item())   # the kernel's punish count
        self._fails_f += fail
        out_info = {"constraints": fail, "success": int(info["success"][0].item())}
        return self._obs_list(obs), self.rewards, self.dones, out_info

    def reset(self, layouts=None):
        self.rewards = {a: 0.0 for a in self.agents}
        self.dones = {a: False for a in self.agents}
        self._fails_f = 0                                         # meda.py:544
        return self._obs_list(self._b.reset(layouts=None if layouts is None else np.asarray(layouts)[None]))

    def restart(self, index=None):
        self.rewards = {a: 0.0 for a in self.agents}
        self.dones = {a: False for a in self.agents}
        obs = self._obs_list(self._b.restart())
        return obs[index] if index else obs                       # meda.py:558-561 (index 0 returns everything)

    def getObs(self):
        return self._obs_list(self._b.get_obs())

    def getOneObs(self, agent_index):
        return self.getObs()[agent_index]

    def render(self, close=False):
        return None

    def seed(self, seed=None):
        pass

    def close(self):
        pass

    def get_env_info(self):
        return self._b.get_env_info()


class MEDAEnv_v0_1(MEDAEnv):
    """env.MEDA.meda.MEDAEnv_v0_1 (meda.py:784-844; `--version 0.1`, common/config.py:14-16): float64 observation
    of length 4*fov^2+2 = all droplets / own goal / goals of the observed others / border, and the direction
    (dy / width, dx / length).  The batched kernel emits int8 with the direction NUMERATORS (dy, dx); this
    adapter divides."""
    _obs_version = nat.MEDA_OBS_V01


class MEDAEnv_v0_2(MEDAEnv):
    """env.MEDA.meda.MEDAEnv_v0_2 (meda.py:846-897): int8 observation of length 3*fov^2+2."""
    _obs_version = nat.MEDA_OBS_V02
