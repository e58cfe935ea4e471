"""Differential tests of the batched CALLERS of the hot path against the reference's own
classes: golden vectors recorded from the unmodified common/rollout.py:RolloutWorker / Evaluator, policy/vdn.py:VDN and
common/replay_buffer.py:ReplayBuffer by tests/golden/make_marl_golden.py, plus live comparisons when the reference
checkout is mounted (build container only).  The CPU tests drive the batched rollout through an oracle-backed env
(tests/oracle_env.py); the GPU tests (-m gpu) drive it through the CUDA envs and the N=1 adapters."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
HAVE_REF = os.path.isdir("/root/reference/policy")


@pytest.fixture(scope="module")
def P():
    return importlib.import_module("marl-dmfb_b200")


# ------------------------------------------------------------------ learner --
def _torch_batch(g, prefix="batch_"):
    return {k[len(prefix):]: torch.as_tensor(v) for k, v in g.items() if k.startswith(prefix)}


def test_vdn_learner_matches_reference_vdn_learn(P):
    """4 consecutive VDN.learn calls of the reference (policy/vdn.py:79-132: double-network TD target, padding mask,
    Adam(0.9, 0.99), grad-norm clip 9, target sync every 2 steps) on a padded 5-episode batch: same weights after every
    call.  A wrong TD target, mask, clip or sync would change them."""
    g = load_golden("marl_vdn_learn")
    lr, gamma, clip, cycle, rnn_dim, hyper = g["hyper"]
    learner = P.VDNLearner((3, 9, 9, 2, 245), 4, 5, "cpu", lr=lr, gamma=gamma, grad_norm_clip=clip,
                           target_update_cycle=int(cycle), rnn_hidden_dim=int(rnn_dim), hyper_hidden_dim=int(hyper))
    init = {k[5:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("init_")}
    learner.eval_rnn.load_state_dict(init)
    learner.target_rnn.load_state_dict(init)
    batch = _torch_batch(g)
    for step in range(4):
        learner.learn({k: v.clone() for k, v in batch.items()}, step)
        for k, v in learner.eval_rnn.state_dict().items():
            np.testing.assert_allclose(v.numpy(), g[f"step{step}_{k}"], rtol=2e-5, atol=2e-7, err_msg=f"step {step} {k}")
        np.testing.assert_allclose(learner.target_rnn.fc1.weight.detach().numpy(), g[f"step{step}_target_fc1.weight"],
                                   rtol=2e-5, atol=2e-7, err_msg=f"target net after step {step}")
    # the target net really was synchronised inside the sequence (after train_step 2) and then left behind again
    assert not np.array_equal(g["step1_target_fc1.weight"], g["step2_target_fc1.weight"])
    assert not np.array_equal(g["step3_fc1.weight"], g["step3_target_fc1.weight"])


def test_learner_accepts_the_batched_format_with_the_same_result(P):
    """The batched EpisodeBatch differs from the reference's wire format in two masked entries (terminal observation
    kept in `o` at the first padded step, avail_u_next zeros at the terminating transition): same loss, same update."""
    g = load_golden("marl_vdn_learn")
    ref_batch = _torch_batch(g)
    mine = {k: v.clone() for k, v in ref_batch.items()}
    live = ~mine["padded"]
    first_pad = mine["padded"] & ~torch.cat([torch.zeros_like(mine["padded"][:, :1]), mine["padded"][:, :-1]], dim=1)
    mine["o"] = torch.where(first_pad[..., None], torch.roll(mine["o_next"], 1, dims=1), mine["o"])
    mine["avail_u_next"] = mine["avail_u_next"] * (live & ~mine["terminated"])[..., None].to(torch.int8)
    outs = []
    for b in (ref_batch, mine):
        learner = P.VDNLearner((3, 9, 9, 2, 245), 4, 5, "cpu", rnn_hidden_dim=16, hyper_hidden_dim=4, seed=3)
        loss = learner.learn(b, 0)
        outs.append((float(loss), torch.cat([p.detach().flatten() for p in learner.eval_rnn.parameters()])))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_vdn_learner_matches_live_reference_at_full_size(P):
    """Same comparison against the live reference at the real network size (290,765 parameters), loss included."""
    sys.path.insert(0, "/root/reference")
    try:
        from policy.vdn import VDN
    finally:
        sys.path.pop(0)
    args = types.SimpleNamespace(
        n_actions=5, n_agents=4, obs_shape=(3, 9, 9, 2, 245), fov=9, last_action=True, reuse_network=True, net="crnn",
        cuda=False, model_dir="/tmp/_no_model", alg="vdn", load_model=False, optimizer="ADAM", lr=5e-4, gamma=0.99,
        grad_norm_clip=9, target_update_cycle=200, rnn_hidden_dim=128, hyper_hidden_dim=24, episode_limit=5)
    torch.manual_seed(4)
    ref = VDN(args)
    mine = P.VDNLearner(args.obs_shape, 4, 5, "cpu")
    mine.eval_rnn.load_state_dict(ref.eval_rnn.state_dict())
    mine.target_rnn.load_state_dict(ref.target_rnn.state_dict())
    rng = np.random.default_rng(1)
    B, T, A, D = 3, 5, 4, 245
    o = rng.integers(0, 5, (B, T + 1, A, D)).astype(np.int8)
    u = rng.integers(0, 5, (B, T, A, 1)).astype(np.int8)
    batch = {"o": o[:, :T].copy(), "o_next": o[:, 1:].copy(), "u": u, "r": rng.normal(size=(B, T, 1)),
             "avail_u": np.ones((B, T, A, 5), np.int8), "avail_u_next": np.ones((B, T, A, 5), np.int8),
             "u_onehot": np.eye(5, dtype=np.int8)[u[..., 0]], "padded": np.zeros((B, T, 1), bool),
             "terminated": np.zeros((B, T, 1), bool)}
    batch["terminated"][:, -1] = True
    batch["terminated"][1, 2:] = True
    batch["padded"][1, 3:] = True
    ref.learn({k: v.copy() for k, v in batch.items()}, T, 0)
    mine.learn({k: torch.as_tensor(v) for k, v in batch.items()}, 0)
    for (k, a), b in zip(ref.eval_rnn.state_dict().items(), mine.eval_rnn.state_dict().values()):
        np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-5, atol=1e-7, err_msg=k)


# ------------------------------------------------------------ replay buffer --
def test_replay_storage_indices_match_reference(P):
    g = load_golden("marl_replay_idx")
    buf = P.ReplayBufferGPU(int(g["size"]), 3, 2, 77, 5, "cpu")
    got = []
    for k, inc in enumerate(g["incs"]):
        got.append(buf._storage_idx(int(inc)).numpy())
        assert buf.current_idx == g["current_idx"][k] and buf.current_size == g["current_size"][k]
    np.testing.assert_array_equal(np.concatenate(got), g["idx"])


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_replay_buffer_contents_match_live_reference(P):
    """store_episode / sample of the reference ReplayBuffer vs ReplayBufferGPU fed with the same episodes: identical
    storage, and a sample drawn with the same indices returns identical arrays for every key."""
    sys.path.insert(0, "/root/reference")
    try:
        from common.replay_buffer import ReplayBuffer
    finally:
        sys.path.pop(0)
    T, A, D, NA = 4, 2, 11, 5
    args = types.SimpleNamespace(n_actions=NA, n_agents=A, obs_shape=(3, 1, 1, 2, D), buffer_size=7, episode_limit=T, alg="vdn")
    ref = ReplayBuffer(args)
    mine = P.ReplayBufferGPU(7, T, A, D, NA, "cpu")
    rng = np.random.default_rng(0)
    for n in (3, 2, 4, 1):
        ep = P.EpisodeBatch(n, T, A, D, NA, "cpu")
        ep.o_all.copy_(torch.as_tensor(rng.integers(0, 5, (T + 1, n, A, D)).astype(np.int8)))
        ep.u.copy_(torch.as_tensor(rng.integers(0, NA, (T, n, A, 1)).astype(np.int8)))
        ep.u_onehot.copy_(torch.nn.functional.one_hot(ep.u[..., 0].long(), NA).to(torch.int8))
        ep.r.copy_(torch.as_tensor(rng.normal(size=(T, n, 1)).astype(np.float32)))
        ep.avail_all.fill_(1)
        ep.padded.fill_(False)
        ep.terminated.fill_(False)
        ep.terminated[-1] = True
        ref.store_episode(ep.to_reference())
        mine.store_episodes(ep)
    assert (ref.current_idx, ref.current_size) == (mine.current_idx, mine.current_size)
    idx = np.array([6, 0, 3, 3, 5])
    got = mine.to_reference(torch.as_tensor(idx))
    for k, v in ref.buffers.items():
        np.testing.assert_array_equal(got[k], v[idx], err_msg=k)


# ------------------------------------------------------------------ rollout --
class ScriptedBatchedAgents:
    """BatchedAgents interface with the actions read from a [N, T, A] script; checks that the worker hands over the
    previous step's one-hot actions (agent.py:29-30) and the current observation."""

    def __init__(self, script, n_actions, device):
        self.script = torch.as_tensor(script.astype(np.int64), device=device)
        self.n_actions, self.t, self.prev = n_actions, 0, None
        self.device = device

    def init_hidden(self, n):
        self.t, self.prev = 0, None
        return torch.zeros(n, 1, device=self.device)

    def choose_actions(self, obs, last_onehot, hidden, avail, epsilon):
        if self.prev is None:
            assert not bool(last_onehot.any())
        else:
            assert torch.equal(last_onehot.to(torch.int8), self.prev)
        a = self.script[:, self.t]
        self.t += 1
        live = avail.any(-1)                                             # padded rows: avail all zero
        self.prev = torch.nn.functional.one_hot(a, self.n_actions).to(torch.int8) * live[..., None]
        return a, hidden


def _check_rollout(P, g, env):
    dev = env.device
    K, T = int(g["K"]), int(g["T"])
    agents = ScriptedBatchedAgents(g["script"], int(g["n_actions"]), dev)
    worker = P.BatchedRolloutWorker(env, agents, epsilon=1.0, min_epsilon=float(g["min_epsilon"]),
                                    anneal_steps=int(g["anneal_steps"]), epsilon_anneal_scale="step", sync_every=0)
    ep, stats = worker.generate_episodes(reset_kwargs={"layouts": g["layouts"].astype(np.uint8)})
    ref = ep.to_reference()
    for key in ("o", "u", "r", "avail_u", "o_next", "avail_u_next", "u_onehot", "terminated", "padded"):
        want = g["ep_" + key]
        if key == "r":
            np.testing.assert_allclose(ref[key], want, rtol=1e-6, atol=1e-7, err_msg=key)
        else:
            np.testing.assert_array_equal(ref[key], want, err_msg=key)
        assert ref[key].dtype == want.dtype, key
    # (reward, steps with failures charged episode_limit, constraints, success) of rollout.py:143-150
    np.testing.assert_allclose(stats["reward"].cpu().numpy(), g["stats"][:, 0], rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(stats["steps"].cpu().numpy(), g["stats"][:, 1])
    cons = stats["constraints"].cpu().numpy().astype(np.float64)
    if str(g["kind"]) == "meda":
        cons = -0.6 * cons                                               # the kernels count punishes (meda.py:326-329)
    np.testing.assert_allclose(cons, g["stats"][:, 2], rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(stats["success"].cpu().numpy(), g["stats"][:, 3])
    # epsilon after the same number of env-steps (rollout.py:126-127); the K episodes ran in lock step here
    assert abs(worker.epsilon - float(g["epsilon"][-1])) < 1e-9
    return ep


def test_batched_rollout_matches_reference_rollout_worker_dmfb(P):
    from oracle_env import OracleBatchedDMFB
    g = load_golden("marl_rollout_c1")
    assert g["stats"][:, 3].sum() >= 2 and (~g["ep_padded"]).sum() < g["ep_padded"].size     # successes and padding occur
    _check_rollout(P, g, OracleBatchedDMFB(int(g["K"]), int(g["W"]), int(g["L"]), int(g["A"]), fov=int(g["fov"])))


def test_batched_rollout_matches_reference_rollout_worker_meda(P):
    from oracle_env import OracleBatchedMEDA
    g = load_golden("marl_rollout_meda")
    _check_rollout(P, g, OracleBatchedMEDA(int(g["K"]), int(g["W"]), int(g["L"]), int(g["A"]), fov=int(g["fov"]), obs_version=0))


def test_epsilon_anneals_per_env_step_like_the_reference(P):
    """with N envs in lock step epsilon must fall by anneal * (live env-steps), not once per iteration."""
    from oracle_env import OracleBatchedDMFB
    g = load_golden("marl_rollout_c1")
    K = int(g["K"])
    env = OracleBatchedDMFB(K, 10, 10, 4, fov=9)
    agents = ScriptedBatchedAgents(g["script"], 5, env.device)
    worker = P.BatchedRolloutWorker(env, agents, epsilon=1.0, min_epsilon=0.05, anneal_steps=100000, sync_every=0)
    _, stats = worker.generate_episodes(reset_kwargs={"layouts": g["layouts"].astype(np.uint8)})
    live_steps = int((~g["ep_padded"]).sum())
    assert abs(worker.epsilon - (1.0 - live_steps * 0.95 / 100000)) < 1e-9


# ---------------------------------------------------------------- GPU tests --
@pytest.mark.gpu
def test_batched_rollout_on_cuda_envs_matches_reference_rollout_worker(P):
    g = load_golden("marl_rollout_c1")
    env = P.BatchedDMFB(int(g["K"]), 10, 10, 4, fov=9, device="cuda:0", layouts=g["layouts"].astype(np.uint8))
    ep = _check_rollout(P, g, env)
    assert ep.o_all.is_cuda
    g = load_golden("marl_rollout_meda")       # config #4: MEDA training env = MEDAEnv (base observation), fov 19
    env = P.BatchedMEDA(int(g["K"]), 30, 60, 4, fov=19, obs_version=0, device="cuda:0", layouts=g["layouts"].astype(np.uint8))
    _check_rollout(P, g, env)


class _ScriptedAgents:
    """The part of agent/agent.py:Agents that common/rollout.py touches, with scripted actions."""

    def __init__(self, n_agents, n_actions):
        self.n_agents, self.n_actions = n_agents, n_actions
        self.policy = types.SimpleNamespace(init_hidden=lambda n: None)
        self.script, self.calls = None, 0

    def start(self, script):
        self.script, self.calls = script, 0

    def choose_action(self, obs, last_action, agent_num, avail_actions, epsilon, evaluate=False):
        inputs = obs.copy()                                    # agent.py:23
        inputs = np.hstack((inputs, last_action))              # agent.py:29-30
        assert inputs.shape == (obs.shape[0] + self.n_actions,)
        t, self.calls = self.calls // self.n_agents, self.calls + 1
        return int(self.script[t][agent_num])


def _evaluator_generate_episode(env, agents, episode_limit):
    """The call sequence of common/rollout.py:Evaluator.one_step / _generate_episode (:19-67), restated (the GPU box has
    no reference checkout): reset() -> per agent choose_action(obs[i], last_action[i], ...) -> step(list of ints) ->
    rewards / dones read per name over env.agents -> render()."""
    n_agents, n_actions = agents.n_agents, agents.n_actions
    obs = env.reset()
    terminated, step, success, reward, constraints = False, 0, 0, 0, 0
    last_action = np.zeros((n_agents, n_actions))
    agents.policy.init_hidden(1)
    while not terminated and step < episode_limit:
        actions = []
        for agent_id in range(n_agents):
            avail_action = [1] * n_actions
            action = agents.choose_action(obs[agent_id], last_action[agent_id], agent_id, avail_action, 0)
            onehot = np.zeros(n_actions)
            onehot[action] = 1
            actions.append(int(action))
            last_action[agent_id] = onehot
        new_obs, r, term, info = env.step(actions)
        r = np.sum([r[agent] for agent in env.agents]) / len(r)
        terminated = np.all([term[agent] for agent in env.agents])
        env.render()
        reward += r
        constraints += info["constraints"]
        success += info["success"]
        obs = new_obs
        step += 1
    if not success:
        step = episode_limit
    return reward, step, constraints, success


@pytest.mark.gpu
def test_n1_adapter_under_the_reference_evaluator_call_sequence(P):
    """Boundary proof: the N=1 adapter DMFBenv driven exactly like common/rollout.py:Evaluator
    drives the reference env - python lists / dicts in and out, obs.copy(), np.hstack with the one-hot, render(),
    close() - reproduces the (reward, steps, constraints, success) the UNMODIFIED Evaluator got from the reference env
    on the same tasks and scripted actions, and Evaluator.evaluate's averages (rollout.py:69-85)."""
    g = load_golden("marl_evaluator_c1")
    K, T, A = int(g["K"]), int(g["T"]), int(g["A"])
    env = P.DMFBenv(int(g["W"]), int(g["L"]), A, fov=int(g["fov"]))
    info = env.get_env_info()
    assert info == {"n_actions": 5, "n_agents": A, "obs_shape": (3, 9, 9, 2, 245), "episode_limit": T}
    agents = _ScriptedAgents(A, 5)
    stats = []
    for k in range(K):
        lay = g["layouts"][k]
        reset = env.reset
        env.reset = lambda lay=lay, reset=reset: reset(layouts=lay)      # the task the reference's generator drew
        agents.start(g["script"][k])
        stats.append(_evaluator_generate_episode(env, agents, T))
        env.reset = reset
    env.close()
    stats = np.array(stats, np.float64)
    np.testing.assert_allclose(stats[:, 0], g["stats"][:, 0], rtol=1e-12, atol=1e-12)     # float64 team rewards
    np.testing.assert_array_equal(stats[:, 1:], g["stats"][:, 1:])
    np.testing.assert_allclose(stats.mean(0), g["mean"], rtol=1e-12)
    assert g["stats"][:, 3].sum() >= 1
