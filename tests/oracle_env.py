"""CPU stand-ins for BatchedDMFB / BatchedMEDA on top of the oracle (TEST INFRASTRUCTURE): the same batched env API
(reset(out=, layouts=) / step(actions, freeze_terminated=, out=) / get_avail_actions / get_env_info) with torch CPU
tensors, so that the host-side callers of the hot path (marl-dmfb_b200/marl.py) can be tested without a GPU.  The
freeze semantics restate DMFB_STEP_FREEZE_TERM (include/dmfb_b200.h): an env whose `terminated` flag is set is not
stepped and emits zero padding (rollout.py:131-141)."""
import numpy as np
import torch

import oracle


class _OracleEnvBase:
    def get_avail_actions(self):
        return self.avail

    def _finish(self, obs, rew, done, cons, succ, frozen, out):
        A = self.A
        obs[frozen] = 0
        rew[frozen] = 0.0
        done[frozen] = 1
        cons[frozen] = 0
        succ[frozen] = 0
        term = done.all(-1)
        self.terminated = term.copy()
        self.avail = torch.as_tensor(np.broadcast_to(np.where(frozen, 0, 1)[:, None, None], (self.N, A, self.n_actions))
                                     .astype(np.uint8).copy())
        o = torch.as_tensor(obs)
        if out is not None:
            out.copy_(o)
            o = out
        info = {"constraints": torch.as_tensor(cons.astype(np.int32)), "success": torch.as_tensor(succ.astype(np.uint8)),
                "terminated": torch.as_tensor(term), "team_reward": torch.as_tensor((rew.sum(-1) / A).astype(np.float32)),
                "padded": torch.as_tensor(frozen.copy())}
        return o, torch.as_tensor(rew.astype(np.float32)), torch.as_tensor(done.astype(bool)), info


class OracleBatchedDMFB(_OracleEnvBase):
    n_actions = 5

    def __init__(self, n_envs, width, length, n_agents, fov=9):
        self.ref = oracle.OracleDMFB(n_envs, width, length, n_agents, fov=fov)
        self.N, self.W, self.L, self.A, self.fov, self.D = n_envs, width, length, n_agents, fov, self.ref.D
        self.max_step = 2 * (width + length)
        self.device = torch.device("cpu")
        self.terminated = np.zeros(n_envs, bool)
        self.avail = torch.ones(n_envs, n_agents, 5, dtype=torch.uint8)

    def get_env_info(self):
        return {"n_actions": 5, "n_agents": self.A, "obs_shape": (3, self.fov, self.fov, 2, self.D), "episode_limit": self.max_step}

    def reset(self, out=None, layouts=None):
        obs = torch.as_tensor(self.ref.reset(np.asarray(layouts)))
        self.terminated[:] = False
        if out is not None:
            out.copy_(obs)
            return out
        return obs

    def step(self, actions, freeze_terminated=False, out=None):
        r = self.ref
        frozen = self.terminated.copy() if freeze_terminated else np.zeros(self.N, bool)
        keep = (r.drop.copy(), r.step_count.copy(), r.constraints.copy(), r.usage.copy())
        obs, rew, done, cons, succ = r.step(actions.cpu().numpy().astype(np.int8))
        for cur, old in zip((r.drop, r.step_count, r.constraints, r.usage), keep):
            cur[frozen] = old[frozen]
        return self._finish(obs, rew, done, cons, succ, frozen, out)

    def get_state(self, out=None):
        s = torch.as_tensor(self.ref.global_state())
        if out is not None:
            out.copy_(s)
            return out
        return s


class OracleBatchedMEDA(_OracleEnvBase):
    n_actions = 9

    def __init__(self, n_envs, width, length, n_agents, fov=19, obs_version=0):
        self.ref = oracle.OracleMEDA(n_envs, width, length, n_agents, fov=fov, obs_version=obs_version)
        self.N, self.W, self.L, self.A, self.fov, self.D = n_envs, width, length, n_agents, fov, self.ref.D
        self.obs_version = obs_version
        self.max_step = width + length
        self.device = torch.device("cpu")
        self.terminated = np.zeros(n_envs, bool)
        self.avail = torch.ones(n_envs, n_agents, 9, dtype=torch.uint8)

    def get_env_info(self):
        c = 3 if self.obs_version == 2 else 4
        return {"n_actions": 9, "n_agents": self.A, "obs_shape": (c, self.fov, self.fov, 2, self.D), "episode_limit": self.max_step}

    def reset(self, out=None, layouts=None):
        obs = torch.as_tensor(self.ref.reset(np.asarray(layouts)))
        self.terminated[:] = False
        if out is not None:
            out.copy_(obs)
            return out
        return obs

    def step(self, actions, freeze_terminated=False, out=None):
        r = self.ref
        frozen = self.terminated.copy() if freeze_terminated else np.zeros(self.N, bool)
        keep = (r.drop.copy(), r.status.copy(), r.step_count.copy(), r.fails.copy(), r.usage.copy())
        obs, rew, done, cons, succ = r.step(actions.cpu().numpy().astype(np.int8))
        for cur, old in zip((r.drop, r.status, r.step_count, r.fails, r.usage), keep):
            cur[frozen] = old[frozen]
        return self._finish(obs, rew, done, cons, succ, frozen, out)
