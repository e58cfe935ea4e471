"""Callers of the hot path, batched: lock-step rollout over N GPU-resident envs with
one batched CRNN forward per step, a device-resident episode / replay buffer in the reference's wire format, and the
VDN learner with a single-bucket gradient all-reduce for data-parallel training.

These are the batched counterparts of
  common/rollout.py:19-39,101-150  (Evaluator.one_step / RolloutWorker.generate_episode)
  agent/agent.py:22-48             (Agents.choose_action: batch-1 CRNN forward per agent, eps-greedy)
  common/replay_buffer.py:5-75     (episode ring buffer, uniform sampling with replacement)
  policy/vdn.py:79-196             (double-network TD learning with the GRU unrolled over the episode)
  network/base_net.py:23-71        (CRNN), network/vdn_net.py (sum mixer)
  policy/qmix.py:73-123, network/qmix_net.py:7-63, network/base_net.py:7-21   (QMIX: monotonic mixer over the global
                                   state; the reference never fills `s`/`s_next` - here they come from the env's
                                   get_state kernel, getglobalobs dmfb.py:368-392)
Plain PyTorch: the networks are tiny dense models (about 0.9 MFLOP per agent-step); the env kernels are the product.
Parameter names follow the reference so that state_dicts interoperate (vdn.py:205-218 file naming kept by
`VDNLearner.save_model`).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------ networks --
def _conv_stack(fov, in_ch, od):
    """network/base_net.py:23-33: conv stack chosen by fov (fov 19 applies the same conv3 module twice)."""
    conv1 = nn.Conv2d(in_ch, od, kernel_size=3, stride=1)
    conv2 = nn.Conv2d(in_ch, od, kernel_size=3, stride=2)
    conv3 = nn.Conv2d(od, od, kernel_size=3, stride=1)
    table = {5: [conv1], 7: [conv1, conv3], 9: [conv1, conv3], 11: [conv1, conv3], 13: [conv1, conv3],
             19: [conv2, conv3, conv3]}
    if fov not in table:
        raise KeyError(fov)  # the reference raises KeyError for any other fov too
    return table[fov]


class CRNN(nn.Module):
    """network/base_net.py:35-71.  obs_shape = (C, fov, fov, 2, C*fov*fov+2); input = obs ++ last action one-hot."""

    def __init__(self, obs_shape, n_actions, rnn_hidden_dim=128, hyper_hidden_dim=24):
        super().__init__()
        self.input_dim = tuple(obs_shape)
        self.rnn_hidden_dim = rnn_hidden_dim
        self.n_actions = n_actions
        fov = obs_shape[1]
        self.convs = _conv_stack(fov, obs_shape[0], hyper_hidden_dim)
        size = fov
        for i, conv in enumerate(self.convs, start=1):
            self.add_module("conv{}".format(i), conv)          # same (shared) registration as base_net.py:47-50
            size = (size + 2 * conv.padding[0] - conv.dilation[0] * (conv.kernel_size[0] - 1) - 1) // conv.stride[0] + 1
        self.out = size * size * hyper_hidden_dim
        self.mlp1 = nn.Linear(obs_shape[-2] + n_actions, 10)
        self.rnn = nn.GRUCell(self.out + 10, rnn_hidden_dim)
        self.fc1 = nn.Linear(rnn_hidden_dim, n_actions)

    def forward(self, inputs, hidden_state, channels_last=False):
        npix = self.input_dim[-1] - self.input_dim[-2]
        pixel, vec = inputs[:, :npix], inputs[:, npix:]
        pixel = pixel.reshape((-1,) + self.input_dim[:3])
        if channels_last:
            pixel = pixel.contiguous(memory_format=torch.channels_last)
        for conv in self.convs:
            pixel = F.relu(conv(pixel))
        pixel = pixel.reshape(-1, self.out)
        vec = F.relu(self.mlp1(vec))
        h = self.rnn(torch.cat([pixel, vec], dim=1), hidden_state.reshape(-1, self.rnn_hidden_dim))
        return self.fc1(h), h


class RNN(nn.Module):
    """network/base_net.py:7-21: the MLP-GRU agent network policy/qmix.py builds (fc1 -> GRUCell -> fc2)."""

    def __init__(self, input_shape, n_actions, rnn_hidden_dim=64):
        super().__init__()
        self.rnn_hidden_dim = rnn_hidden_dim
        self.n_actions = n_actions
        self.fc1 = nn.Linear(input_shape, rnn_hidden_dim)
        self.rnn = nn.GRUCell(rnn_hidden_dim, rnn_hidden_dim)
        self.fc2 = nn.Linear(rnn_hidden_dim, n_actions)

    def forward(self, obs, hidden_state):
        x = F.relu(self.fc1(obs))
        h = self.rnn(x, hidden_state.reshape(-1, self.rnn_hidden_dim))
        return self.fc2(h), h


class QMixNet(nn.Module):
    """network/qmix_net.py:7-63: hypernetworks of the global state produce the non-negative mixing weights."""

    def __init__(self, state_shape, n_agents, qmix_hidden_dim=32, hyper_hidden_dim=64, two_hyper_layers=False):
        super().__init__()
        self.state_shape, self.n_agents, self.qmix_hidden_dim = state_shape, n_agents, qmix_hidden_dim
        if two_hyper_layers:
            self.hyper_w1 = nn.Sequential(nn.Linear(state_shape, hyper_hidden_dim), nn.ReLU(),
                                          nn.Linear(hyper_hidden_dim, n_agents * qmix_hidden_dim))
            self.hyper_w2 = nn.Sequential(nn.Linear(state_shape, hyper_hidden_dim), nn.ReLU(),
                                          nn.Linear(hyper_hidden_dim, qmix_hidden_dim))
        else:
            self.hyper_w1 = nn.Linear(state_shape, n_agents * qmix_hidden_dim)
            self.hyper_w2 = nn.Linear(state_shape, qmix_hidden_dim)
        self.hyper_b1 = nn.Linear(state_shape, qmix_hidden_dim)
        self.hyper_b2 = nn.Sequential(nn.Linear(state_shape, qmix_hidden_dim), nn.ReLU(), nn.Linear(qmix_hidden_dim, 1))

    def forward(self, q_values, states):
        """q_values [B, T, n_agents], states [B, T, state_shape] -> q_total [B, T, 1]."""
        B = q_values.size(0)
        q_values = q_values.reshape(-1, 1, self.n_agents)
        states = states.reshape(-1, self.state_shape)
        w1 = torch.abs(self.hyper_w1(states)).view(-1, self.n_agents, self.qmix_hidden_dim)
        b1 = self.hyper_b1(states).view(-1, 1, self.qmix_hidden_dim)
        hidden = F.elu(torch.bmm(q_values, w1) + b1)
        w2 = torch.abs(self.hyper_w2(states)).view(-1, self.qmix_hidden_dim, 1)
        b2 = self.hyper_b2(states).view(-1, 1, 1)
        return (torch.bmm(hidden, w2) + b2).view(B, -1, 1)


# ------------------------------------------------------------------- timing --
class PhaseTimer:
    """CUDA-event stopwatch per named phase (env step / policy forward / learner / all-reduce), on the current stream.
    Nothing synchronises until summary(); with enabled=False every call is a no-op."""

    def __init__(self, enabled=True):
        self.enabled, self.pairs = enabled, {}

    class _Span:
        def __init__(self, timer, name):
            self.t, self.name = timer, name

        def __enter__(self):
            if self.t.enabled:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *a):
            if self.t.enabled:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                self.t.pairs.setdefault(self.name, []).append((self.e0, e1))

    def __call__(self, name):
        return PhaseTimer._Span(self, name)

    def summary(self, reset=True):
        """{phase: total milliseconds since the last summary}"""
        if not self.enabled:
            return {}
        torch.cuda.synchronize()
        out = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.pairs.items()}
        if reset:
            self.pairs = {}
        return out


_NO_TIMER = PhaseTimer(enabled=False)


# ------------------------------------------------------------ action selection --
class BatchedAgents:
    """Agents.choose_action (agent/agent.py:22-48) for all N*A agents in one forward pass, on the device."""

    def __init__(self, net, n_agents, n_actions, device, seed=0, autocast_dtype=None):
        # autocast_dtype (e.g. torch.bfloat16): the ROLLOUT's forward pass runs under torch.autocast with channels-last
        # activations - an opt-in trade of the reference's fp32 action values for throughput (the learner stays fp32)
        self.net, self.n_agents, self.n_actions = net, n_agents, n_actions
        self.autocast_dtype = autocast_dtype
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))

    def init_hidden(self, n_envs):
        return torch.zeros(n_envs * self.n_agents, self.net.rnn_hidden_dim, device=self.device)

    @torch.no_grad()
    def choose_actions(self, obs, last_onehot, hidden, avail, epsilon):
        """obs int8 [N,A,D], last_onehot [N,A,n_actions], hidden [N*A,H], avail [N,A,n_actions] (0/1).
        Returns actions int64 [N,A] and the new hidden state.  Padded rows (avail all zero) get action 0."""
        N, A = obs.shape[0], obs.shape[1]
        if self.autocast_dtype is None:
            inputs = torch.cat([obs.to(torch.float32), last_onehot.to(torch.float32)], dim=2).reshape(N * A, -1)
            q, hidden = self.net(inputs, hidden)
        else:
            dt = self.autocast_dtype
            inputs = torch.cat([obs.to(dt), last_onehot.to(dt)], dim=2).reshape(N * A, -1)
            with torch.autocast("cuda", dtype=dt):
                q, hidden = self.net(inputs, hidden.to(dt), channels_last=True)
            q, hidden = q.float(), hidden.float()
        q = q.reshape(N, A, self.n_actions).masked_fill(avail == 0, float("-inf"))     # agent.py:43
        greedy = torch.argmax(torch.nan_to_num(q, neginf=-3.0e38), dim=2)
        # np.random.choice(avail_actions_ind) (agent.py:44-45): uniform over the available actions
        w = avail.to(torch.float32) + (avail.sum(dim=2, keepdim=True) == 0).to(torch.float32)
        rand = torch.multinomial(w.reshape(N * A, -1), 1, generator=self.gen).reshape(N, A)
        explore = torch.rand(N, A, device=self.device, generator=self.gen) < epsilon
        return torch.where(explore, rand, greedy), hidden


# --------------------------------------------------------------- episode batch --
class EpisodeBatch:
    """n lock-step episodes of T = episode_limit transitions, on the device, stored TIME-MAJOR: `o_all[t]` is one
    contiguous [n, A, D] slice, so the env step kernel writes the observation of step t straight into the buffer
    (`env.step(out=ep.o_all[t + 1])`) and the learner's per-transition reads are contiguous too.  `o` and `avail_u` are
    stored for T+1 steps: o_next[t] = o[t+1], avail_u_next[t] = avail_u[t+1].

    as_dict()       the keys and [B, T, ...] shapes of the reference's episode dict (common/replay_buffer.py:17-26),
                    as transposed VIEWS of the time-major storage.  Loss-equivalent, not byte-identical, to what the
                    reference stores: at the first padded step `o` still shows the terminal observation (reference:
                    zeros) and at the terminating transition `avail_u_next` is zeros (reference: ones); both entries
                    are multiplied by the padding mask / by (1 - terminated) in the loss (vdn.py:105-118).
    to_reference()  the reference's wire format bit for bit (rollout.py:101-150 incl. the padding rules :131-141 and
                    the dtypes of replay_buffer.py:17-26), materialised as numpy arrays - for exchanging episodes
                    with the reference's own ReplayBuffer / learner and for the differential tests."""

    KEYS = ("o", "u", "r", "avail_u", "avail_u_next", "u_onehot", "padded", "terminated")
    _FIELDS = ("o_all", "u", "r", "avail_all", "u_onehot", "padded", "terminated")

    def __init__(self, n, T, A, D, n_actions, device, state_dim=0):
        z = lambda *s, dtype: torch.zeros(*s, dtype=dtype, device=device)  # noqa: E731
        self.n, self.T, self.A, self.D, self.n_actions = n, T, A, D, n_actions
        # QMIX only: global state for T+1 steps (s_next[t] = s[t+1]), the flattened (3,W,L) get_state tensor
        self.state_dim = state_dim
        self.s_all = z(T + 1, n, state_dim, dtype=torch.int8) if state_dim else None
        self.o_all = z(T + 1, n, A, D, dtype=torch.int8)
        self.u = z(T, n, A, 1, dtype=torch.int8)
        self.r = z(T, n, 1, dtype=torch.float32)
        self.avail_all = z(T + 1, n, A, n_actions, dtype=torch.int8)
        self.u_onehot = z(T, n, A, n_actions, dtype=torch.int8)
        self.padded = torch.ones(T, n, 1, dtype=torch.bool, device=device)
        self.terminated = torch.ones(T, n, 1, dtype=torch.bool, device=device)

    def as_dict(self, idx=None, T=None):
        T = self.T if T is None else T
        s = (lambda x: x) if idx is None else (lambda x: x[:, idx])          # episodes live on dim 1
        v = lambda x, a, b: s(x)[a:b].transpose(0, 1)                        # noqa: E731  -> [B, T, ...] view
        d = {"o": v(self.o_all, 0, T), "o_next": v(self.o_all, 1, T + 1), "u": v(self.u, 0, T), "r": v(self.r, 0, T),
             "avail_u": v(self.avail_all, 0, T), "avail_u_next": v(self.avail_all, 1, T + 1),
             "u_onehot": v(self.u_onehot, 0, T), "padded": v(self.padded, 0, T), "terminated": v(self.terminated, 0, T)}
        if self.s_all is not None:
            d["s"], d["s_next"] = v(self.s_all, 0, T), v(self.s_all, 1, T + 1)
        return d

    def to_reference(self, idx=None):
        d = {k: v.cpu().numpy() for k, v in self.as_dict(idx).items()}
        live = ~d["padded"]                                                   # [B, T, 1]
        out = {"o": (d["o"] * live[..., None]).astype("int8"), "u": d["u"].astype("int8"),
               "r": d["r"].astype("float64"), "o_next": d["o_next"].astype("int8"),
               "avail_u": d["avail_u"].astype("int8"),
               "avail_u_next": (live[..., None] * (d["avail_u"] * 0 + 1)).astype("int8"),   # ones on every live step
               "u_onehot": d["u_onehot"].astype("int8"), "padded": d["padded"].astype(bool),
               "terminated": d["terminated"].astype(bool)}
        if "s" in d:
            out["s"], out["s_next"] = d["s"], d["s_next"]
        return out


class ReplayBufferGPU(EpisodeBatch):
    """common/replay_buffer.py:5-75 on the device: ring of `size` episodes, uniform sampling with replacement."""

    def __init__(self, size, T, A, D, n_actions, device, seed=0, state_dim=0):
        super().__init__(size, T, A, D, n_actions, device, state_dim=state_dim)
        self.size, self.current_idx, self.current_size = size, 0, 0
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(int(seed))

    def _storage_idx(self, inc):
        """_get_storage_idx (replay_buffer.py:58-75)."""
        dev = self.o_all.device
        if self.current_idx + inc <= self.size:
            idx = torch.arange(self.current_idx, self.current_idx + inc, device=dev)
            self.current_idx += inc
        elif self.current_idx < self.size:
            overflow = inc - (self.size - self.current_idx)
            idx = torch.cat([torch.arange(self.current_idx, self.size, device=dev), torch.arange(0, overflow, device=dev)])
            self.current_idx = overflow
        else:
            idx = torch.arange(0, inc, device=dev)
            self.current_idx = inc
        self.current_size = min(self.size, self.current_size + inc)
        return idx

    def store_episodes(self, ep):
        if ep.n > self.size:
            raise ValueError("more episodes than the buffer holds")
        idx = self._storage_idx(ep.n)
        for name in self._FIELDS:
            getattr(self, name).index_copy_(1, idx, getattr(ep, name))
        if self.s_all is not None:
            self.s_all.index_copy_(1, idx, ep.s_all)
        return idx

    def sample(self, batch_size):
        idx = torch.randint(0, self.current_size, (batch_size,), device=self.o_all.device, generator=self.gen)
        return self.as_dict(idx)


# -------------------------------------------------------------------- rollout --
class BatchedRolloutWorker:
    """RolloutWorker.generate_episode (rollout.py:101-150) for all N envs of a batched env in lock step.

    Finished envs are frozen by the env kernel (`freeze_terminated`), which emits exactly the zero padding the
    reference appends (rollout.py:131-141); the loop stops as soon as every env is done.  The step kernel writes each
    observation directly into the episode buffer slice `ep.o_all[t + 1]`.

    Epsilon: the reference anneals once per ENV step (rollout.py:126-127).  Here all N envs of a lock-step iteration act
    under the same epsilon, which is then annealed by the number of env-steps that iteration really made (the live
    envs), with the reference's stopping rule (no further decrement once epsilon <= min_epsilon) - i.e. after the same
    total number of env-steps epsilon is where the reference's would be.  It lives on the device, so annealing costs no
    host synchronisation; `self.epsilon` is read back once per rollout."""

    def __init__(self, env, agents, epsilon=1.0, min_epsilon=0.05, anneal_steps=150000, epsilon_anneal_scale="step",
                 record_state=False, sync_every=8, timer=None):
        self.env, self.agents = env, agents
        self.timer = timer or _NO_TIMER
        self.record_state = record_state      # QMIX: store get_state() of every step as `s` / `s_next`
        info = env.get_env_info()
        self.T, self.A, self.n_actions, self.D = info["episode_limit"], info["n_agents"], info["n_actions"], info["obs_shape"][-1]
        self.epsilon, self.min_epsilon = epsilon, min_epsilon
        self.anneal_epsilon = (epsilon - min_epsilon) / anneal_steps
        self.epsilon_anneal_scale = epsilon_anneal_scale
        self.sync_every = sync_every

    def _anneal(self, eps, count):
        """`count` applications of `eps = eps - anneal if eps > min_epsilon else eps` (rollout.py:115,127), closed form."""
        if self.anneal_epsilon <= 0:
            return eps
        # decrements until eps <= min_epsilon; a quotient within 1e-6 of an integer is that integer (the reference gets
        # there by repeated float subtraction, which lands a few ulps on either side of min_epsilon)
        k = torch.clamp(torch.ceil((eps - self.min_epsilon) / self.anneal_epsilon - 1e-6), min=0)
        return eps - torch.minimum(k, count.to(eps.dtype)) * self.anneal_epsilon

    @torch.no_grad()
    def generate_episodes(self, evaluate=False, batch=None, reset_kwargs=None):
        """Returns (EpisodeBatch, stats) with stats = per-env episode reward, steps (episode_limit when not successful,
        rollout.py:148-149), constraints and success, all device tensors.  `reset_kwargs` go to env.reset (e.g.
        injected `layouts`)."""
        env, N, A, T = self.env, self.env.N, self.A, self.T
        dev = env.device
        state_dim = 3 * env.W * env.L if self.record_state else 0
        ep = batch if batch is not None else EpisodeBatch(N, T, A, self.D, self.n_actions, dev, state_dim=state_dim)
        ep.padded.fill_(True); ep.terminated.fill_(True)
        ep.u.zero_(); ep.u_onehot.zero_(); ep.r.zero_(); ep.avail_all.zero_(); ep.o_all[1:].zero_()
        env.reset(out=ep.o_all[0], **(reset_kwargs or {}))
        if ep.s_all is not None:
            ep.s_all[1:].zero_()
            env.get_state(out=ep.s_all[0].view(N, 3, env.W, env.L))
        ep.avail_all[0] = 1
        hidden = self.agents.init_hidden(N)
        last = torch.zeros(N, A, self.n_actions, device=dev)
        reward = torch.zeros(N, device=dev)
        constraints = torch.zeros(N, device=dev, dtype=torch.int64)
        success = torch.zeros(N, device=dev, dtype=torch.int64)
        steps = torch.zeros(N, device=dev, dtype=torch.int64)
        alive = torch.ones(N, dtype=torch.bool, device=dev)
        eps = torch.full((), 0.0 if evaluate else float(self.epsilon), device=dev, dtype=torch.float64)
        if not evaluate and self.epsilon_anneal_scale == "episode":
            eps = self._anneal(eps, torch.tensor(N, device=dev))
        for t in range(T):
            with self.timer("policy_forward"):
                actions, hidden = self.agents.choose_actions(ep.o_all[t], last, hidden, ep.avail_all[t], eps)
            with self.timer("env_step"):
                _, _, _, info = env.step(actions, freeze_terminated=True, out=ep.o_all[t + 1])   # written in place
            live = alive & ~info["padded"]
            onehot = F.one_hot(actions, self.n_actions).to(torch.int8) * live[:, None, None]
            if ep.s_all is not None:                            # padded steps keep the zero state, like the zero obs
                env.get_state(out=ep.s_all[t + 1].view(N, 3, env.W, env.L))
                ep.s_all[t + 1] *= live[:, None]
            ep.u[t, :, :, 0] = (actions * live[:, None]).to(torch.int8)
            ep.u_onehot[t] = onehot
            ep.r[t, :, 0] = info["team_reward"]
            ep.avail_all[t + 1] = env.get_avail_actions() * (live & ~info["terminated"])[:, None, None]
            ep.padded[t, :, 0] = ~live
            ep.terminated[t, :, 0] = info["terminated"] | ~live
            reward += info["team_reward"] * live
            constraints += info["constraints"].to(torch.int64) * live
            success += info["success"].to(torch.int64) * live
            steps += live.to(torch.int64)
            last = onehot.to(torch.float32)
            alive = live & ~info["terminated"]
            if not evaluate and self.epsilon_anneal_scale == "step":
                eps = self._anneal(eps, live.sum())
            if self.sync_every and t % self.sync_every == self.sync_every - 1 and not bool(alive.any()):
                break                                        # one host sync every `sync_every` steps
        if not evaluate:
            self.epsilon = float(eps)
        steps = torch.where(success > 0, steps, torch.full_like(steps, T))
        return ep, {"reward": reward, "steps": steps, "constraints": constraints, "success": (success > 0).to(torch.int64)}


# -------------------------------------------------------------------- learner --
def global_mask_sum(mask_sum, world_size):
    """Number of valid transitions over ALL ranks.  Each rank scales its loss by world_size / global count, so that the
    rank-averaged gradient is the gradient of the mean over every valid transition of the global batch ("data parallel
    == one large batch" also when the ranks hold different numbers of valid transitions)."""
    if world_size <= 1:
        return mask_sum
    import torch.distributed as dist
    total = mask_sum.detach().clone()
    dist.all_reduce(total)
    return total


def allreduce_gradients(params, world_size):
    """One flat fp32 bucket, summed over ranks and divided by the world size (290,765 parameters =
    1.16 MB per learn step; latency-bound on NVLink, so a single bucket)."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or world_size <= 1:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat)
    flat.div_(world_size)
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)


class VDNLearner:
    """policy/vdn.py: eval/target CRNN, sum mixer, Adam(0.9, 0.99), grad-norm clip, periodic target sync."""

    def __init__(self, obs_shape, n_agents, n_actions, device, lr=5e-4, gamma=0.99, grad_norm_clip=9.0,
                 target_update_cycle=200, rnn_hidden_dim=128, hyper_hidden_dim=24, world_size=1, seed=0, timer=None):
        torch.manual_seed(seed)
        self.timer = timer or _NO_TIMER
        self.device = torch.device(device)
        self.n_agents, self.n_actions = n_agents, n_actions
        self.eval_rnn = CRNN(obs_shape, n_actions, rnn_hidden_dim, hyper_hidden_dim).to(self.device)
        self.target_rnn = CRNN(obs_shape, n_actions, rnn_hidden_dim, hyper_hidden_dim).to(self.device)
        self.target_rnn.load_state_dict(self.eval_rnn.state_dict())
        self.eval_parameters = list(self.eval_rnn.parameters())             # VDNNet has no parameters (vdn_net.py:5-10)
        self.optimizer = torch.optim.Adam(self.eval_parameters, lr=lr, betas=(0.9, 0.99))   # vdn.py:66-68
        self.gamma, self.grad_norm_clip, self.target_update_cycle = gamma, grad_norm_clip, target_update_cycle
        self.world_size = world_size

    @staticmethod
    def max_episode_len(batch):
        """Agents._get_max_episode_len (agent.py:51-61): 1 + the latest first-terminated index over the batch."""
        term = batch["terminated"][:, :, 0]
        T = term.shape[1]
        first = torch.where(term.any(dim=1), term.to(torch.int64).argmax(dim=1) + 1, torch.zeros_like(term[:, 0], dtype=torch.int64))
        return int(min(T, first.max().item()))

    def _inputs(self, batch, t):
        """VDN._get_inputs (vdn.py:134-165): obs of step t (o for t = 0, o_next[t-1] afterwards) ++ previous action."""
        B = batch["o"].shape[0]
        obs = batch["o"][:, 0] if t == 0 else batch["o_next"][:, t - 1]
        prev = torch.zeros_like(batch["u_onehot"][:, 0]) if t == 0 else batch["u_onehot"][:, t - 1]
        return torch.cat([obs.to(torch.float32), prev.to(torch.float32)], dim=2).reshape(B * self.n_agents, -1)

    def q_values(self, batch, T):
        """VDN.get_q_values (vdn.py:167-196): GRU unrolled over the T transitions for both networks."""
        B = batch["o"].shape[0]
        h_e = torch.zeros(B * self.n_agents, self.eval_rnn.rnn_hidden_dim, device=self.device)
        h_t = torch.zeros_like(h_e)
        q_evals, q_targets = [], []
        inputs = self._inputs(batch, 0)
        for t in range(T):
            inputs_next = self._inputs(batch, t + 1) if t + 1 <= batch["o_next"].shape[1] else inputs
            q_e, h_e = self.eval_rnn(inputs, h_e)
            with torch.no_grad():
                q_t, h_t = self.target_rnn(inputs_next, h_t)
            q_evals.append(q_e.view(B, self.n_agents, -1))
            q_targets.append(q_t.view(B, self.n_agents, -1))
            inputs = inputs_next
        return torch.stack(q_evals, dim=1), torch.stack(q_targets, dim=1)

    def learn(self, batch, train_step, max_episode_len=None):
        """VDN.learn (vdn.py:79-132).  batch: dict of device tensors [B, T, ...] (EpisodeBatch.as_dict / sample)."""
        T = self.max_episode_len(batch) if max_episode_len is None else max_episode_len
        T = max(T, 1)
        batch = {k: v[:, :T] for k, v in batch.items()}                      # agent.py:66-68
        u = batch["u"].to(torch.int64)
        r = batch["r"].to(torch.float32)
        terminated = batch["terminated"].to(torch.float32)
        mask = 1.0 - batch["padded"].to(torch.float32)
        q_evals, q_targets = self.q_values(batch, T)
        q_evals = torch.gather(q_evals, dim=3, index=u).squeeze(3)
        q_targets = q_targets.masked_fill(batch["avail_u_next"] == 0, -9999999.0).max(dim=3)[0]
        q_total_eval = q_evals.sum(dim=2, keepdim=True)                      # VDNNet
        q_total_target = q_targets.sum(dim=2, keepdim=True)
        targets = r + self.gamma * q_total_target * (1.0 - terminated)
        masked_td = mask * (targets.detach() - q_total_eval)
        denom = global_mask_sum(mask.sum(), self.world_size).clamp_min(1.0)
        loss = (masked_td ** 2).sum() * (float(self.world_size) / denom)     # world_size 1: sum / mask.sum() (vdn.py:116)
        self.optimizer.zero_grad(set_to_none=False)
        loss.backward()
        with self.timer("grad_allreduce"):
            allreduce_gradients(self.eval_parameters, self.world_size)
        torch.nn.utils.clip_grad_norm_(self.eval_parameters, self.grad_norm_clip)
        self.optimizer.step()
        if train_step > 0 and train_step % self.target_update_cycle == 0:
            self.target_rnn.load_state_dict(self.eval_rnn.state_dict())
        return loss.detach()

    def save_model(self, model_dir, ith_run=0, train_step=None):
        """File naming of VDN.save_model (vdn.py:205-218) so that the reference's evaluate.py can load the weights."""
        os.makedirs(model_dir, exist_ok=True)
        mid = "" if train_step is None else str(train_step) + "_"
        torch.save({}, os.path.join(model_dir, "{}_{}vdn_net_params.pkl".format(ith_run, mid)))
        torch.save(self.eval_rnn.state_dict(), os.path.join(model_dir, "{}_{}rnn_net_params.pkl".format(ith_run, mid)))


class QMIXLearner(VDNLearner):
    """policy/qmix.py:73-123: the VDN learner with the sum replaced by QMixNet(q, s) / QMixNet_target(q', s_next).
    The reference builds its agents from the MLP-GRU `RNN` and never produces `s`; here the agent network is the same
    CRNN the batched rollout uses (so `BatchedAgents` / `BatchedRolloutWorker(record_state=True)` feed it unchanged)
    and `s` is the env's global state (3*W*L int8, get_state)."""

    def __init__(self, obs_shape, n_agents, n_actions, state_shape, device, qmix_hidden_dim=32, hyper_hidden_dim_mix=64,
                 two_hyper_layers=False, **kw):
        super().__init__(obs_shape, n_agents, n_actions, device, **kw)
        self.eval_qmix_net = QMixNet(state_shape, n_agents, qmix_hidden_dim, hyper_hidden_dim_mix, two_hyper_layers).to(self.device)
        self.target_qmix_net = QMixNet(state_shape, n_agents, qmix_hidden_dim, hyper_hidden_dim_mix, two_hyper_layers).to(self.device)
        self.target_qmix_net.load_state_dict(self.eval_qmix_net.state_dict())
        self.eval_parameters = list(self.eval_qmix_net.parameters()) + list(self.eval_rnn.parameters())   # qmix.py:53-54
        lr = kw.get("lr", 5e-4)
        self.optimizer = torch.optim.Adam(self.eval_parameters, lr=lr, betas=(0.9, 0.99))                   # qmix.py:61-63

    def learn(self, batch, train_step, max_episode_len=None):
        T = self.max_episode_len(batch) if max_episode_len is None else max_episode_len
        T = max(T, 1)
        batch = {k: v[:, :T] for k, v in batch.items()}
        u = batch["u"].to(torch.int64)
        r = batch["r"].to(torch.float32)
        s, s_next = batch["s"].to(torch.float32), batch["s_next"].to(torch.float32)
        terminated = batch["terminated"].to(torch.float32)
        mask = 1.0 - batch["padded"].to(torch.float32)
        q_evals, q_targets = self.q_values(batch, T)
        q_evals = torch.gather(q_evals, dim=3, index=u).squeeze(3)
        q_targets = q_targets.masked_fill(batch["avail_u_next"] == 0, -9999999.0).max(dim=3)[0]
        q_total_eval = self.eval_qmix_net(q_evals, s)
        with torch.no_grad():
            q_total_target = self.target_qmix_net(q_targets, s_next)
        targets = r + self.gamma * q_total_target * (1.0 - terminated)
        masked_td = mask * (q_total_eval - targets.detach())
        denom = global_mask_sum(mask.sum(), self.world_size).clamp_min(1.0)
        loss = (masked_td ** 2).sum() * (float(self.world_size) / denom)
        self.optimizer.zero_grad(set_to_none=False)
        loss.backward()
        allreduce_gradients(self.eval_parameters, self.world_size)
        torch.nn.utils.clip_grad_norm_(self.eval_parameters, self.grad_norm_clip)
        self.optimizer.step()
        if train_step > 0 and train_step % self.target_update_cycle == 0:
            self.target_rnn.load_state_dict(self.eval_rnn.state_dict())
            self.target_qmix_net.load_state_dict(self.eval_qmix_net.state_dict())
        return loss.detach()

    def save_model(self, model_dir, ith_run=0, train_step=None):
        """File naming of QMIX.save_model / save_final_model (qmix.py:196-212)."""
        os.makedirs(model_dir, exist_ok=True)
        mid = "" if train_step is None else str(train_step) + "_"
        torch.save(self.eval_qmix_net.state_dict(), os.path.join(model_dir, "{}_{}qmix_net_params.pkl".format(ith_run, mid)))
        torch.save(self.eval_rnn.state_dict(), os.path.join(model_dir, "{}_{}rnn_net_params.pkl".format(ith_run, mid)))

