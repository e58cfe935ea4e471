#!/usr/bin/env python
"""Golden vectors for the CALLERS of the hot path, recorded from the UNMODIFIED reference classes.

Run in the build container only (needs /root/reference):

    python tests/golden/make_marl_golden.py

 marl_vdn_learn.npz   policy/vdn.py:VDN - initial weights, one padded episode batch, and the weights / losses after every
                      one of 4 `learn` calls (Adam, grad-norm clip, a target sync in between)   -> VDNLearner.learn
 marl_rollout_*.npz   common/rollout.py:RolloutWorker.generate_episode on the reference env under a scripted policy:
                      layouts, scripted actions, the episode dicts the reference would store in its ReplayBuffer, the
                      (reward, step, constraints, success) tuples and the epsilon after every episode
                                                                                          -> BatchedRolloutWorker
 marl_evaluator_c1.npz common/rollout.py:Evaluator._generate_episode / evaluate under the same kind of script
                                                                                          -> N=1 adapter (boundary proof)
 marl_replay_idx.npz  common/replay_buffer.py:ReplayBuffer._get_storage_idx over a sequence of batch sizes
                                                                                          -> ReplayBufferGPU._storage_idx
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_dmfb, ref_meda = ref_shim.install()
from common.replay_buffer import ReplayBuffer  # noqa: E402  (reference)
from common.rollout import Evaluator, RolloutWorker  # noqa: E402  (reference)
from policy.vdn import VDN  # noqa: E402  (reference)


def vdn_args(**kw):
    a = types.SimpleNamespace(
        n_actions=5, n_agents=4, obs_shape=(3, 9, 9, 2, 245), fov=9, last_action=True, reuse_network=True, net="crnn",
        cuda=False, model_dir="/tmp/_no_model", alg="vdn", load_model=False, optimizer="ADAM", lr=5e-4, gamma=0.99,
        grad_norm_clip=9, target_update_cycle=2, rnn_hidden_dim=16, hyper_hidden_dim=4, episode_limit=6)
    a.__dict__.update(kw)
    return a


def gen_vdn_learn():
    torch.manual_seed(11)
    args = vdn_args()
    pol = VDN(args)
    init = {k: v.detach().clone().numpy() for k, v in pol.eval_rnn.state_dict().items()}
    rng = np.random.default_rng(5)
    B, T, A, D, NA = 5, args.episode_limit, args.n_agents, 245, 5
    lens = [6, 3, 5, 1, 4]                                   # live transitions per episode
    o = rng.integers(0, 5, (B, T + 1, A, D)).astype(np.int8)
    u = rng.integers(0, NA, (B, T, A, 1)).astype(np.int8)
    batch = {"o": o[:, :T].copy(), "o_next": o[:, 1:].copy(), "u": u, "r": rng.normal(size=(B, T, 1)),
             "avail_u": np.ones((B, T, A, NA), np.int8), "avail_u_next": np.ones((B, T, A, NA), np.int8),
             "u_onehot": np.eye(NA, dtype=np.int8)[u[..., 0]], "padded": np.zeros((B, T, 1), bool),
             "terminated": np.zeros((B, T, 1), bool)}
    for b, n in enumerate(lens):                             # the padding rules of rollout.py:131-141
        batch["terminated"][b, n - 1:] = True
        batch["padded"][b, n:] = True
        for k in ("o", "o_next", "u", "r", "avail_u", "avail_u_next", "u_onehot"):
            batch[k][b, n:] = 0
    out = {"init_" + k: v for k, v in init.items()}
    out.update({"batch_" + k: v for k, v in batch.items()})
    losses = []
    for step in range(4):
        b = {k: v.copy() for k, v in batch.items()}
        # Agents.train (agent.py:63-68): truncate to the longest episode of the batch
        max_len = 0
        for e in range(B):
            for t in range(T):
                if b["terminated"][e, t, 0] == 1:
                    max_len = max(max_len, t + 1)
                    break
        for k in b:
            b[k] = b[k][:, :max_len]
        pol.learn(b, max_len, step)
        for k, v in pol.eval_rnn.state_dict().items():
            out[f"step{step}_{k}"] = v.detach().clone().numpy()
        out[f"step{step}_target_fc1.weight"] = pol.target_rnn.state_dict()["fc1.weight"].detach().clone().numpy()
        # the loss is not returned by the reference: recompute it with the updated nets only for the record
        losses.append(0.0)
    out["hyper"] = np.array([args.lr, args.gamma, args.grad_norm_clip, args.target_update_cycle, args.rnn_hidden_dim,
                             args.hyper_hidden_dim], np.float64)
    np.savez_compressed(os.path.join(HERE, "marl_vdn_learn.npz"), **out)
    print("marl_vdn_learn: %d arrays" % len(out))


class ScriptedAgents:
    """Stands in for agent/agent.py:Agents: the interface Evaluator / RolloutWorker use (n_agents, n_actions,
    choose_action, policy.init_hidden) with the actions read from a script.  It also performs the input handling of
    Agents.choose_action (agent.py:23-30: obs.copy(), np.hstack with the last action) so that the env's return types
    are exercised exactly like the real class does."""

    def __init__(self, n_agents, n_actions):
        self.n_agents, self.n_actions = n_agents, n_actions
        self.policy = types.SimpleNamespace(init_hidden=lambda n: None)
        self.script, self.calls, self.input_lens = None, 0, set()

    def start(self, script):
        self.script, self.calls = script, 0

    def choose_action(self, obs, last_action, agent_num, avail_actions, epsilon, evaluate=False):
        inputs = obs.copy()
        inputs = np.hstack((inputs, last_action))
        self.input_lens.add(inputs.shape)
        t, self.calls = self.calls // self.n_agents, self.calls + 1
        assert self.calls % self.n_agents == (agent_num + 1) % self.n_agents
        return int(self.script[t][agent_num])


def gen_rollout(name, kind, K, seed, p_goal, **kw):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    import random as pyrandom
    if kind == "dmfb":
        env = ref_dmfb.DMFBenv(kw["W"], kw["L"], kw["A"], 0, fov=kw["fov"])
        n_actions = 5
    else:
        env = ref_meda.MEDAEnv(kw["W"], kw["L"], kw["A"], fov=kw["fov"])     # config('meda', '0.2') -> MEDAEnv
        n_actions = 9
    pyrandom.seed(seed)
    info = env.get_env_info()
    if kind == "meda":
        # the reference's MEDA get_env_info returns the flat length (meda.py:676-681) although RolloutWorker / CRNN index
        # a tuple (rollout.py:94, base_net.py:38-40): as shipped `train.py meda` cannot start.  The tuple a working
        # configuration needs is the DMFB-style one.
        assert info["obs_shape"] == 4 * kw["fov"] ** 2 + 2
        info["obs_shape"] = (4, kw["fov"], kw["fov"], 2, info["obs_shape"])
    A, T, D = info["n_agents"], info["episode_limit"], info["obs_shape"][-1]
    args = types.SimpleNamespace(episode_limit=T, n_actions=n_actions, obs_shape=info["obs_shape"],
                                 epsilon_anneal_scale="step", epsilon=1.0, min_epsilon=0.05, anneal_steps=120)
    agents = ScriptedAgents(A, n_actions)
    worker = RolloutWorker(env, agents, args)
    # The script must not depend on the env's reaction (the batched worker replays it blindly), but should still drive
    # droplets home: plan it on the task the reset is GOING to draw.  reset() draws the task from the global RNGs, so
    # capture their state, peek at the task with a throw-away reset, restore, and let generate_episode redraw it.
    out = dict(kind=kind, K=K, A=A, T=T, D=D, n_actions=n_actions, W=kw["W"], L=kw["L"], fov=kw["fov"],
               layouts=np.zeros((K, A, 4), np.int16), script=np.zeros((K, T, A), np.int8),
               stats=np.zeros((K, 4), np.float64), epsilon=np.zeros(K, np.float64),
               anneal_steps=args.anneal_steps, min_epsilon=args.min_epsilon)
    eps_keys = ("o", "u", "r", "avail_u", "o_next", "avail_u_next", "u_onehot", "terminated", "padded")
    episodes = {k: [] for k in eps_keys}
    for k in range(K):
        st_np, st_py = np.random.get_state(), pyrandom.getstate()
        env.reset()
        rm = env.routing_manager
        if kind == "dmfb":
            lay = [(d.x, d.y, d.des_x, d.des_y) for d in rm.droplets]
        else:
            lay = [(d.x_center, d.y_center, g.x_center, g.y_center) for d, g in zip(rm.droplets, rm.destinations)]
        np.random.set_state(st_np)
        pyrandom.setstate(st_py)
        out["layouts"][k] = lay
        # open-loop plan: walk each droplet along x then y towards its goal (ignoring the others), noise with 1 - p_goal
        script = np.zeros((T, A), np.int8)
        pos = [[c[0], c[1]] for c in lay]
        for t in range(T):
            for i in range(A):
                dx, dy = lay[i][2] - pos[i][0], lay[i][3] - pos[i][1]
                if rng.random() < p_goal and (dx or dy):
                    if kind == "dmfb":
                        a = (1 if dx > 0 else 2) if dx else (4 if dy > 0 else 3)
                        pos[i][0] += (a == 1) - (a == 2)
                        pos[i][1] += (a == 4) - (a == 3)
                    else:
                        if abs(dx) >= 3 or (dx and not dy):
                            a = 1 if dx > 0 else 3
                            pos[i][0] += 3 if dx > 0 else -3
                        elif dy:
                            a = 2 if dy > 0 else 0
                            pos[i][1] += 3 if dy > 0 else -3
                        else:
                            a = 8
                else:
                    a = int(rng.integers(n_actions))
                script[t, i] = a
        if kind == "dmfb" and k % 2 == 0:
            # every other episode: a closed-loop, constraint-avoiding plan (make_golden._dmfb_safe_policy) recorded on
            # the peeked task, so that some episodes END IN SUCCESS; the env is deterministic here (health == 1), so
            # replaying the recorded actions after the RNG restore reproduces the same trajectory
            import make_golden
            st2_np, st2_py = np.random.get_state(), pyrandom.getstate()
            env.reset()
            for t in range(T):
                acts = make_golden._dmfb_safe_policy(rng, env, 1.0)
                script[t] = acts
                env.step(acts)
            np.random.set_state(st2_np)
            pyrandom.setstate(st2_py)
        out["script"][k] = script
        agents.start(script)
        reward, step, constraints, success, episode = worker.generate_episode()
        out["stats"][k] = (reward, step, constraints, success)
        out["epsilon"][k] = worker.epsilon
        for key in eps_keys:
            episodes[key].append(episode[key][0])
    # what the reference's ReplayBuffer would hold (dtypes of replay_buffer.py:17-26)
    dt = {"o": np.int8, "u": np.int8, "r": np.float64, "o_next": np.int8, "avail_u": np.int8, "avail_u_next": np.int8,
          "u_onehot": np.int8, "padded": bool, "terminated": bool}
    for key in eps_keys:
        arr = np.stack(episodes[key])
        if key in ("o", "o_next"):
            assert np.all(arr == np.round(arr)) and np.abs(arr).max() < 128
        out["ep_" + key] = arr.astype(dt[key])
    assert agents.input_lens == {(D + n_actions,)}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  steps={out['stats'][:, 1].tolist()} "
          f"success={out['stats'][:, 3].tolist()} eps_end={out['epsilon'][-1]:.4f}")


def gen_evaluator():
    rng = np.random.default_rng(77)
    np.random.seed(77)
    W, L, A, fov, K = 10, 10, 4, 9, 10
    env = ref_dmfb.DMFBenv(W, L, A, 0, fov=fov)
    info = env.get_env_info()
    T = info["episode_limit"]
    agents = ScriptedAgents(A, 5)
    ev = Evaluator(env, agents, T)
    out = dict(W=W, L=L, A=A, fov=fov, K=K, T=T, layouts=np.zeros((K, A, 4), np.int16), script=np.zeros((K, T, A), np.int8),
               stats=np.zeros((K, 4), np.float64))
    for k in range(K):
        st = np.random.get_state()
        env.reset()
        lay = [(d.x, d.y, d.des_x, d.des_y) for d in env.routing_manager.droplets]
        np.random.set_state(st)
        out["layouts"][k] = lay
        script = np.zeros((T, A), np.int8)
        pos = [[c[0], c[1]] for c in lay]
        for t in range(T):
            for i in range(A):
                dx, dy = lay[i][2] - pos[i][0], lay[i][3] - pos[i][1]
                if rng.random() < 0.9 and (dx or dy):
                    a = (1 if dx > 0 else 2) if dx else (4 if dy > 0 else 3)
                    pos[i][0] += (a == 1) - (a == 2)
                    pos[i][1] += (a == 4) - (a == 3)
                else:
                    a = int(rng.integers(5))
                script[t, i] = a
        out["script"][k] = script
        agents.start(script)
        out["stats"][k] = ev._generate_episode()
    out["mean"] = out["stats"].mean(0)            # what Evaluator.evaluate(K) returns (rollout.py:69-85)
    np.savez_compressed(os.path.join(HERE, "marl_evaluator_c1.npz"), **out)
    print("marl_evaluator_c1: stats mean", out["mean"].tolist())


def gen_replay_idx():
    rng = np.random.default_rng(3)
    args = types.SimpleNamespace(n_actions=5, n_agents=2, obs_shape=(3, 5, 5, 2, 77), buffer_size=23, episode_limit=3, alg="vdn")
    buf = ReplayBuffer(args)
    incs = rng.integers(1, 12, 40)
    idx, cur, size = [], [], []
    for inc in incs:
        i = buf._get_storage_idx(int(inc))
        idx.append(np.atleast_1d(i))
        cur.append(buf.current_idx)
        size.append(buf.current_size)
    np.savez_compressed(os.path.join(HERE, "marl_replay_idx.npz"), size=23, incs=incs, idx=np.concatenate(idx),
                        current_idx=np.array(cur), current_size=np.array(size))
    print("marl_replay_idx: ok")


if __name__ == "__main__":
    gen_vdn_learn()
    gen_rollout("marl_rollout_c1", "dmfb", K=10, seed=31, p_goal=0.85, W=10, L=10, A=4, fov=9)
    gen_rollout("marl_rollout_meda", "meda", K=4, seed=32, p_goal=0.85, W=30, L=60, A=4, fov=19)
    gen_evaluator()
    gen_replay_idx()
