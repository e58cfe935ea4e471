"""Host-side mirror of the reference DMFB env interface (env/DMFB/dmfb.py) over GPU-resident state.

* ``BatchedDMFB`` — N independent chips in HBM, stepped in lock step by the sm_100a kernels through
  the C ABI (include/dmfb_b200.h).  Tensors in, tensors out, no host synchronisation inside ``step``.
* ``DMFBenv``     — N = 1 adapter returning the reference's exact Python types (lists of np.int8
  arrays, {"player_i": float} dicts, info dict) so that the reference's own
  ``common/rollout.py:Evaluator`` / ``RolloutWorker`` can drive it unmodified.

PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as nat
from .pipeline import SubBatches, offset_ptr


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedDMFB:
    """N chips x A droplets.  Mirrors DMFBenv (dmfb.py:474-640) with a leading env dimension.

    reset(mask=None, new=False, layouts=None, degrade=None) -> obs[N,A,D] int8
    step(actions[N,A]) -> obs, reward[N,A] f32, done[N,A] bool, info{constraints, success, terminated, team_reward}
    get_obs() / get_state() / get_avail_actions() / get_env_info()
    """

    n_actions = 5

    def __init__(self, n_envs, width, length, n_agents, n_blocks=0, fov=5, stall=True, b_degrade=False,
                 per_degrade=0.1, device="cuda", seed=0, env_base=0, track_usage=None, reward_f64=False,
                 degrade=None, layouts=None, block_layouts=None, obs_version=0, usage_log=True, task_prefetch=True,
                 health_bitmap=True, sub_batches=1, search_period=None):
        # sub_batches: K > 1 steps the batch as K sub-batches on K streams (pipeline.py): consecutive steps issued with
        # step(..., join=False) then overlap - the results are the same, env for env
        self.lib = nat.load()
        self.cfg = nat.DmfbCfg()
        nat.check(self.lib.dmfb_cfg_init(C.byref(self.cfg), width, length, n_agents, n_blocks, fov, int(bool(stall)),
                                         int(bool(b_degrade)), float(per_degrade)), "dmfb_cfg_init")
        # obs_version 1 = DMFBenv_v0_1.getOneObs (dmfb.py:723-835): 4 layers; the two direction entries hold the
        # integer numerators (tar_y - y, tar_x - x) of the reference's ((tar_y - y) / length, (tar_x - x) / width)
        self.obs_version = int(obs_version)
        if self.obs_version:
            nat.check(self.lib.dmfb_cfg_set_obs_version(C.byref(self.cfg), self.obs_version), "dmfb_cfg_set_obs_version")
        if n_blocks and self.cfg.n_blocks == 0:
            print('Too many required modules in the environment.')    # dmfb.py:232-234: continues without blocks
        self.n_blocks = int(self.cfg.n_blocks)
        self.cfg.env_base = int(env_base)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedDMFB needs a CUDA device: there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.N, self.W, self.L, self.A, self.fov = int(n_envs), width, length, n_agents, fov
        self.D = self.cfg.obs_dim
        self.max_step = self.cfg.max_step
        self.stall, self.b_degrade, self.per_degrade = bool(stall), bool(b_degrade), float(per_degrade)
        self.seed = int(seed)
        self.agents = ["player_{}".format(i) for i in range(n_agents)]
        N, A, dev = self.N, self.A, self.device
        if track_usage is None:
            track_usage = self.b_degrade
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        # ---- state (struct of arrays over N) ----
        self.drop = z(N, A, 4, dtype=torch.uint8)
        self.start = z(N, A, 2, dtype=torch.uint8)
        self.step_count = z(N, dtype=torch.int32)
        self.constraints_cum = z(N, dtype=torch.int32)
        self.terminated = z(N, dtype=torch.uint8)
        self.episode = z(N, dtype=torch.int32)
        self.usage = z(N, width, length, dtype=torch.int32) if track_usage else None
        self.blocks = z(N, self.n_blocks, 2, dtype=torch.uint8) if self.n_blocks else None   # (x_min, y_min) of 2x2 blocks
        self._health = torch.ones(N, width, length, dtype=torch.float64, device=dev) if self.b_degrade else None
        # degraded-cell bit map (dmfb_state_t.health_bits): lets the step skip the float64 gather on healthy cells.
        # Kernels keep it in step with `health`; `env.health` hands the tensor out and may be written through, so every
        # access marks the map stale and the next call rebuilds it first.
        self._health_bits = (z(N, (width * length + 31) // 32, dtype=torch.int32)
                             if (self.b_degrade and health_bitmap) else None)
        self._health_dirty = False
        self.degrade = torch.ones(N, width, length, dtype=torch.float64, device=dev) if self.b_degrade else None
        # Steps append the actuated cells to a per-env log instead of incrementing `usage` in place; resets (and
        # usage_counts()) fold the log in.  `usage` alone is therefore NOT m_usage between resets: read usage_counts().
        self._usage_log = bool(usage_log) and track_usage
        self.usage_log = z(N, self.max_step, A, dtype=torch.int16) if self._usage_log else None
        self.usage_log_len = z(N, dtype=torch.int32) if self._usage_log else None
        # task prefetch for auto_reset (dmfb_state_t.next_task): the search for the next episode's task runs ahead, a
        # bounded number of attempts per step; the tasks drawn are the same with or without it
        self._prefetch = bool(task_prefetch)
        # dense 10-droplet chips: the run-ahead search is a kernel behind the step (dmfb_step).  Launched behind every
        # P-th step of a (sub-)batch only, with P times the attempts - sub-batch k in the steps with (step + k) % P == 0,
        # so that a step of the whole batch carries about one search launch - it costs one extra kernel boundary per
        # P steps instead of one per step; the tasks drawn are the same
        if search_period is None:
            search_period = int(os.environ.get("DMFB_SEARCH_PERIOD", "4"))
        self._search_period = max(1, min(int(search_period), 255))
        self._auto_steps = 0
        self.next_task = z(N, A, dtype=torch.int32) if self._prefetch else None
        self.next_cursor = z(N, dtype=torch.int32) if self._prefetch else None
        self.status = z(1, dtype=torch.int32)          # sticky DMFB_STATUS_* bits (illegal action, sampler gave up)
        self._track_usage = bool(track_usage)
        self.state = self._make_state(0, N)
        # ---- per-step outputs ----
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
        self._sub = SubBatches(self.device, N, sub_batches) if int(sub_batches) > 1 else None
        if self._sub is not None:
            self._sub_cfg, self._sub_state = [], []
            for lo, hi in self._sub.ranges:
                cfg = nat.DmfbCfg.from_buffer_copy(self.cfg)
                cfg.env_base = self.cfg.env_base + lo            # RNG streams are functions of the GLOBAL env index
                self._sub_cfg.append(cfg)
                self._sub_state.append(self._make_state(lo, hi))
            self._sub_out = [self._make_out(self.obs, lo) for lo, _ in self._sub.ranges]
        # the reference constructor draws the degradation matrix and a first task (dmfb.py:151-155)
        self.reset(new=True, layouts=layouts, degrade=degrade, block_layouts=block_layouts)

    # ------------------------------------------------------------------ helpers --
    def _make_state(self, lo, hi):
        """dmfb_state_t over the envs [lo, hi) of the batch (pointers offset into the same tensors)."""
        p = offset_ptr
        return nat.DmfbState(
            n_envs=hi - lo, usage_log_cap=self.max_step if self._usage_log else 0,
            drop=p(self.drop, lo), start=p(self.start, lo), step_count=p(self.step_count, lo),
            constraints=p(self.constraints_cum, lo), terminated=p(self.terminated, lo), episode=p(self.episode, lo),
            usage=p(self.usage, lo) if self._track_usage else None, health=p(self._health, lo),
            degrade=p(self.degrade, lo), blocks=p(self.blocks, lo), usage_log=p(self.usage_log, lo),
            usage_log_len=p(self.usage_log_len, lo), next_task=p(self.next_task, lo), next_cursor=p(self.next_cursor, lo),
            gen_status=self.status.data_ptr(), health_bits=p(self._health_bits, lo))

    def _make_out(self, obs, lo=0):
        p = offset_ptr
        return nat.DmfbOut(
            obs=p(obs, lo), reward=p(self.reward, lo), reward_f64=p(self.reward_f64, lo),
            team_reward=p(self.team_reward, lo), done=p(self.done, lo), avail=p(self.avail, lo),
            constraints=p(self.constraints, lo), success=p(self.success, lo),
            terminated=p(self.term_out, lo), padded=p(self.padded, lo), status=self.status.data_ptr())

    def join(self):
        """Makes the caller's stream wait for sub-batch steps issued with join=False (no-op otherwise)."""
        if self._sub is not None:
            self._sub.join()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def health(self):
        """m_health [N,W,L] float64 (None without degradation).  The tensor may be written through; the degraded-cell
        bit map is rebuilt before the next kernel that reads it."""
        if self._health_bits is not None:
            self._health_dirty = True
        return self._health

    def _sync_health(self):
        if self._health_dirty:
            self.join()
            self._health_dirty = False
            with torch.cuda.device(self.device):
                rc = self.lib.dmfb_sync_health_bits(C.byref(self.cfg), C.byref(self.state), self._stream())
            nat.check(rc, "dmfb_sync_health_bits")

    def _as(self, t, dtype, shape, name):
        if t is None:
            return None
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.ascontiguousarray(t))
        t = t.to(device=self.device, dtype=dtype).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    # ---------------------------------------------------------------------- API --
    def reset(self, mask=None, new=False, layouts=None, degrade=None, out=None, block_layouts=None):
        """DMFBenv.reset(new) (dmfb.py:589-597) for the envs selected by ``mask`` (None = all).

        layouts: optional [N,A,4] (x, y, goal_x, goal_y) injected tasks; default: on-device generator
        equivalent to _Generate_Start_End (dmfb.py:207-226).  degrade: optional [N,W,L] float64 factors
        (only used with new=True)."""
        self.join()
        mask_t = self._as(mask, torch.uint8, (self.N,), "mask")
        lay_t = self._as(layouts, torch.uint8, (self.N, self.A, 4), "layouts")
        deg_t = self._as(degrade, torch.float64, (self.N, self.W, self.L), "degrade")
        blk_t = self._as(block_layouts, torch.uint8, (self.N, self.n_blocks, 2), "block_layouts") if self.n_blocks else None
        obs = self.obs if out is None else out
        self._sync_health()
        with torch.cuda.device(self.device):
            rc = self.lib.dmfb_reset(C.byref(self.cfg), C.byref(self.state), _ptr(mask_t), int(bool(new)), _ptr(lay_t),
                                     _ptr(blk_t), _ptr(deg_t), self.seed, _ptr(obs), self._stream())
        nat.check(rc, "dmfb_reset")
        return obs

    def restart(self, mask=None):
        """DMFBenv.restart (dmfb.py:599-605)."""
        self.join()
        mask_t = self._as(mask, torch.uint8, (self.N,), "mask")
        with torch.cuda.device(self.device):
            rc = self.lib.dmfb_restart(C.byref(self.cfg), C.byref(self.state), _ptr(mask_t), _ptr(self.obs),
                                       self._stream())
        nat.check(rc, "dmfb_restart")
        return self.obs

    def step(self, actions, draws=None, record=True, freeze_terminated=False, auto_reset=False, out=None, join=True):
        """DMFBenv.step (dmfb.py:560-587) on every env.

        actions: [N,A] integer tensor (int8 / int32 / int64) on this device.
        draws:   optional [N,A] float64 move-success draws replacing random.random() (dmfb.py:335).
        freeze_terminated: lock-step episodes — finished envs are not stepped and emit zero padding.
        auto_reset: envs that terminate in this step are reset (new task, updateHealth) right after it;
                 their obs rows hold the first observation of the next episode.
        out:     optional int8 [N,A,D] tensor that receives the observation (e.g. a slice of an
                 episode buffer) instead of ``self.obs``.
        join:    with sub_batches > 1 only.  False leaves the sub-batch streams running, so that the next step's
                 kernels overlap this one's; the returned tensors are then valid on the caller's stream after join()
                 (every other method joins first).  Keep `actions` / `draws` alive until then."""
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.ascontiguousarray(actions))
        if actions.device != self.device:
            actions = actions.to(self.device)
        if actions.dtype not in (torch.int8, torch.uint8, torch.int32, torch.int64):
            actions = actions.to(torch.int64)
        actions = actions.contiguous()
        if tuple(actions.shape) != (self.N, self.A):
            raise RuntimeError("The number of actions is not the same as n_droplets")  # dmfb.py:272-274
        draws_t = self._as(draws, torch.float64, (self.N, self.A), "draws")
        flags = ((nat.STEP_RECORD_USAGE if record else 0) | (nat.STEP_FREEZE_TERM if freeze_terminated else 0)
                 | (nat.STEP_AUTO_RESET if auto_reset else 0))
        if out is None:
            obs, o = self.obs, self._out
        else:
            obs, o = out, self._make_out(out)
        self._sync_health()
        P, t = self._search_period, self._auto_steps
        if auto_reset:
            self._auto_steps += 1

        def search_flags(k):
            if not auto_reset or P == 1:
                return 0
            return nat.step_search_share(P) if (t + k) % P == 0 else nat.STEP_SKIP_TASK_SEARCH

        if self._sub is None:
            with torch.cuda.device(self.device):
                rc = self.lib.dmfb_step(C.byref(self.cfg), C.byref(self.state), _ptr(actions), actions.element_size(),
                                        _ptr(draws_t), self.seed, flags | search_flags(0), C.byref(o), self._stream())
            nat.check(rc, "dmfb_step")
        else:
            sub, es = self._sub, actions.element_size()
            sub.fork(keep_alive=(actions, draws_t, out))
            with torch.cuda.device(self.device):
                for k, ((lo, hi), stream) in enumerate(zip(sub.ranges, sub.handles())):
                    o_k = self._sub_out[k] if out is None else self._make_out(out, lo)
                    rc = self.lib.dmfb_step(C.byref(self._sub_cfg[k]), C.byref(self._sub_state[k]),
                                            C.c_void_p(actions.data_ptr() + lo * self.A * es), es,
                                            None if draws_t is None else C.c_void_p(draws_t.data_ptr() + lo * self.A * 8),
                                            self.seed, flags | search_flags(k), C.byref(o_k), stream)
                    nat.check(rc, "dmfb_step")
            if join:
                sub.join()
        info = {"constraints": self.constraints, "success": self.success, "terminated": self.term_out.view(torch.bool),
                "team_reward": self.team_reward, "padded": self.padded.view(torch.bool)}
        return obs, self.reward, self.done.view(torch.bool), info

    def get_obs(self, out=None):
        """getObs() of the current state (dmfb.py:622-626), recomputed from the droplet positions."""
        self.join()
        obs = self.obs if out is None else out
        with torch.cuda.device(self.device):
            rc = self.lib.dmfb_observe(C.byref(self.cfg), C.byref(self.state), _ptr(obs), self._stream())
        nat.check(rc, "dmfb_observe")
        return obs

    def get_state(self, out=None):
        """getglobalobs() (dmfb.py:368-392) as int8 [N,3,W,L]."""
        self.join()
        if out is None:
            out = torch.empty(self.N, 3, self.W, self.L, dtype=torch.int8, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.dmfb_global_state(C.byref(self.cfg), C.byref(self.state), _ptr(out), self._stream())
        nat.check(rc, "dmfb_global_state")
        return out

    def get_avail_actions(self):
        """[N,A,n_actions] mask: all ones (rollout.py:22), zeros for padded steps (rollout.py:138-139)."""
        return self.avail

    def get_env_info(self):
        """dmfb.py:633-640."""
        return {"n_actions": self.n_actions, "n_agents": self.A,
                "obs_shape": (4 if self.obs_version else 3, self.fov, self.fov, 2, self.D),
                "episode_limit": self.max_step}

    def check_actions(self):
        """Raises the reference's TypeError if an illegal action was applied since the last check (dmfb.py:115-116),
        and RuntimeError if a task / obstacle generator gave up on a density that cannot be placed (the reference
        would loop for ever there, dmfb.py:212-224,246-250; the env kept its previous layout).  Device -> host sync."""
        self.join()
        st = int(self.status.item())
        if st:
            self.status.zero_()
        if st & nat.STATUS_ILLEGAL_ACTION:
            raise TypeError("action is illegal")
        if st & nat.STATUS_SAMPLER_GAVE_UP:
            raise RuntimeError("the task generator found no legal layout for this chip size / droplet count / obstacles")

    check = check_actions

    # convenience views
    @property
    def positions(self):
        return self.drop[:, :, 0:2]

    @property
    def goals(self):
        return self.drop[:, :, 2:4]

    def usage_counts(self):
        """m_usage as an integer tensor (folds the usage log into the counters first)."""
        self.join()
        if self._usage_log:
            with torch.cuda.device(self.device):
                rc = self.lib.dmfb_flush_usage(C.byref(self.cfg), C.byref(self.state), self._stream())
            nat.check(rc, "dmfb_flush_usage")
        return self.usage


class _RoutingManagerView:
    """The attributes of RoutingTaskManager that callers read (evaDegre.py:21; dmfb.py:194-195)."""

    def __init__(self, env):
        self._e = env

    @property
    def m_health(self):
        b = self._e._b
        return np.ones((b.W, b.L)) if b.health is None else b.health[0].cpu().numpy()

    @property
    def m_degrade(self):
        b = self._e._b
        return np.ones((b.W, b.L)) if b.degrade is None else b.degrade[0].cpu().numpy()

    @property
    def m_usage(self):
        b = self._e._b
        return np.zeros((b.W, b.L)) if b.usage is None else b.usage_counts()[0].cpu().numpy().astype(np.float64)

    @property
    def starts(self):
        return self._e._b.start[0].cpu().numpy().astype(int)

    @property
    def ends(self):
        return self._e._b.drop[0, :, 2:4].cpu().numpy().astype(int)

    @property
    def distances(self):
        d = self._e._b.drop[0].cpu().numpy().astype(int)
        return np.abs(d[:, 0] - d[:, 2]) + np.abs(d[:, 1] - d[:, 3])

    def getTaskStatus(self):
        return [bool(x == 0) for x in self.distances]


class DMFBenv:
    """Drop-in for env.DMFB.dmfb.DMFBenv (dmfb.py:474-640) on top of a 1-env GPU batch.

    Same constructor signature, same return types: reset() -> list of np.int8 arrays of length
    3*fov*fov+2; step(list|dict) -> (obs list, rewards dict, dones dict, info dict)."""

    metadata = {"render.modes": ["human", "rgb_array"]}
    _obs_version = nat.DMFB_OBS_BASE

    def __init__(self, width, length, n_agents, n_blocks=0, fov=5, stall=True, b_degrade=False, per_degrade=0.1,
                 show=False, savemp4=False, device="cuda", seed=None, layouts=None, degrade=None, block_layouts=None):
        assert width >= 5 and length >= 5
        assert n_agents > 0
        if seed is None:
            seed = int(np.random.randint(0, 2**31 - 1))  # the reference seeds from the wall clock (dmfb.py:154)
        self._b = BatchedDMFB(1, width, length, n_agents, n_blocks, fov=fov, stall=stall, b_degrade=b_degrade,
                              per_degrade=per_degrade, device=device, seed=seed, track_usage=True, reward_f64=True,
                              layouts=None if layouts is None else np.asarray(layouts)[None],
                              degrade=None if degrade is None else np.asarray(degrade)[None],
                              block_layouts=None if block_layouts is None else np.asarray(block_layouts)[None],
                              obs_version=self._obs_version)
        self.mode = None  # rendering is out of scope (dmfb.py:642-720)
        self.agents = list(self._b.agents)
        self.possible_agents = self.agents[:]
        self.width, self.length = width, length
        self.max_step = self._b.max_step
        self.rewards = {a: 0.0 for a in self.agents}
        self.dones = {a: False for a in self.agents}
        self.routing_manager = _RoutingManagerView(self)

    @property
    def step_count(self):
        return int(self._b.step_count[0].item())

    @property
    def constraints(self):
        return int(self._b.constraints_cum[0].item())

    def _obs_list(self, obs):
        o = obs[0].cpu().numpy()
        return [o[i].copy() for i in range(len(self.agents))]

    def step(self, actions, record=True):
        if isinstance(actions, dict):
            acts = [actions[a] for a in self.agents]
        elif isinstance(actions, list):
            acts = actions
        else:
            raise TypeError("wrong actions")                     # dmfb.py:563-568
        if len(acts) != len(self.agents):
            raise RuntimeError("The number of actions is not the same as n_droplets")
        a = torch.as_tensor(np.asarray([int(x) for x in acts], dtype=np.int64)[None])
        obs, _, done, info = self._b.step(a, record=record)
        self._b.check_actions()
        r = self._b.reward_f64[0].cpu().numpy()
        d = done[0].cpu().numpy()
        for k, name in enumerate(self.agents):
            self.rewards[name] = r[k]
            self.dones[name] = bool(d[k])
        out_info = {"constraints": int(info["constraints"][0].item()), "success": int(info["success"][0].item())}
        return self._obs_list(obs), self.rewards, self.dones, out_info

    def reset(self, new=False, layouts=None):
        self.rewards = {a: 0 for a in self.agents}
        self.dones = {a: False for a in self.agents}
        obs = self._b.reset(new=new, layouts=None if layouts is None else np.asarray(layouts)[None])
        return self._obs_list(obs)

    def restart(self, index=None):
        self.rewards = {a: 0.0 for a in self.agents}
        self.dones = {a: False for a in self.agents}
        return self._obs_list(self._b.restart())

    def seed(self, seed=None):
        pass                                                     # dmfb.py:607-608

    def close(self):
        pass

    def render(self, close=False):
        return None                                              # mode is None (dmfb.py:643-644)

    def getOneObs(self, agent):
        index = int(agent[-1]) if isinstance(agent, str) else agent  # dmfb.py:614-620
        return self._obs_list(self._b.get_obs())[index]

    def getObs(self):
        return self._obs_list(self._b.get_obs())

    def get_env_info(self):
        return self._b.get_env_info()


class DMFBenv_v0_1(DMFBenv):
    """env.DMFB.dmfb.DMFBenv_v0_1 (dmfb.py:723-835; `--version 0.1`, common/config.py:6-8): float64 observation of
    length 4*fov^2+2 = droplets / own goal / goals of the visible others drawn where the ray to the goal leaves
    the window / obstacles + border, then ((tar_y - y) / length, (tar_x - x) / width).  The batched kernel emits
    int8 with the direction NUMERATORS; this adapter divides."""
    _obs_version = nat.DMFB_OBS_V01

    def _obs_list(self, obs):
        o = obs[0].cpu().numpy().astype(np.float64)
        o[:, -2] = o[:, -2] / self.length                        # dmfb.py:833
        o[:, -1] = o[:, -1] / self.width
        return [o[i].copy() for i in range(len(self.agents))]

    def get_env_info(self):
        # the reference takes obs_shape from RoutingTaskManager.getOneObs, i.e. from the BASE observation
        # (dmfb.py:633-640), also for this subclass
        info = self._b.get_env_info()
        f = self._b.fov
        info["obs_shape"] = (3, f, f, 2, 3 * f * f + 2)
        return info

