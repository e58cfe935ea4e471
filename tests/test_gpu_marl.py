"""GPU checks of the batched callers on top of the CUDA env: lock-step rollout bookkeeping
follows rollout.py:101-150, the episode batch has the reference wire format and padding, and a few VDN updates run."""
import importlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_lockstep_rollout_and_vdn_updates():
    P = importlib.import_module("marl-dmfb_b200")
    dev = torch.device("cuda:0")
    N = 512
    env = P.BatchedDMFB(N, 10, 10, 4, fov=9, device=dev, seed=7)
    info = env.get_env_info()
    T, A, D, n_act = info["episode_limit"], 4, 245, 5
    learner = P.VDNLearner(info["obs_shape"], A, n_act, dev, seed=0)
    agents = P.BatchedAgents(learner.eval_rnn, A, n_act, dev, seed=1)
    worker = P.BatchedRolloutWorker(env, agents, epsilon=1.0, anneal_steps=1000)
    ep, stats = worker.generate_episodes()
    b = ep.as_dict()
    assert b["o"].shape == (N, T, A, D) and b["o"].dtype == torch.int8 and b["u"].shape == (N, T, A, 1)
    assert b["r"].shape == (N, T, 1) and b["avail_u"].shape == (N, T, A, n_act) and b["padded"].dtype == torch.bool
    pad = b["padded"][:, :, 0]
    # padding is a suffix, and padded transitions are all zero / terminated = 1 (rollout.py:131-141)
    assert bool((pad[:, 1:] >= pad[:, :-1]).all()) and not bool(pad[:, 0].any())
    assert not bool(b["o_next"][pad].any()) and not bool(b["u_onehot"][pad].any()) and not bool(b["r"][pad].any())
    assert bool(b["terminated"][:, :, 0][pad].all()) and not bool(b["avail_u"][pad].any())
    # exactly one non-padded terminated transition per episode, at its last live step
    live_term = b["terminated"][:, :, 0] & ~pad
    assert bool((live_term.sum(1) == 1).all())
    n_live = (~pad).sum(1)
    assert bool((live_term.to(torch.int64).argmax(1) + 1 == n_live).all())
    # one-hot matches u on live transitions; avail all ones there
    live = ~pad
    assert bool((b["u_onehot"].argmax(-1)[live] == b["u"][..., 0][live]).all()) and bool((b["avail_u"][live] == 1).all())
    # stats: failures are charged episode_limit steps (rollout.py:148-149)
    ok = stats["success"] > 0
    assert bool((stats["steps"][~ok] == T).all()) and bool((stats["steps"][ok] == n_live[ok]).all())
    np.testing.assert_allclose(stats["reward"].cpu().numpy(), b["r"][:, :, 0].sum(1).cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert worker.epsilon < 1.0                                    # annealed per step (rollout.py:126-127)
    # replay + learning
    buf = P.ReplayBufferGPU(1024, T, A, D, n_act, dev, seed=3)
    buf.store_episodes(ep)
    w0 = learner.eval_rnn.fc1.weight.detach().clone()
    losses = [float(learner.learn(buf.sample(64), s)) for s in range(5)]
    assert np.isfinite(losses).all() and not torch.equal(w0, learner.eval_rnn.fc1.weight)
    # greedy evaluation run (epsilon 0) leaves epsilon untouched
    e0 = worker.epsilon
    _, st = worker.generate_episodes(evaluate=True)
    assert worker.epsilon == e0 and st["steps"].shape == (N,)


def test_qmix_rollout_records_the_global_state_and_learns():
    """QMIX (policy/qmix.py:73-123) on top of the env's get_state kernel: `s` of a live step is getglobalobs() of
    that step, padded steps carry zeros, and a few updates move both the agent network and the mixer."""
    P = importlib.import_module("marl-dmfb_b200")
    dev = torch.device("cuda:0")
    N, W, L, A, n_act = 256, 10, 10, 4, 5
    env = P.BatchedDMFB(N, W, L, A, fov=9, device=dev, seed=11)
    info = env.get_env_info()
    T = info["episode_limit"]
    learner = P.QMIXLearner(info["obs_shape"], A, n_act, 3 * W * L, dev, seed=0)
    agents = P.BatchedAgents(learner.eval_rnn, A, n_act, dev, seed=1)
    worker = P.BatchedRolloutWorker(env, agents, epsilon=1.0, anneal_steps=1000, record_state=True)
    ep, stats = worker.generate_episodes()
    b = ep.as_dict()
    assert b["s"].shape == (N, T, 3 * W * L) and b["s"].dtype == torch.int8
    pad = b["padded"][:, :, 0]
    assert not bool(b["s_next"][pad].any())
    s0 = b["s"][:, 0].reshape(N, 3, W, L)
    # layer 0 of getglobalobs marks every droplet with idx+1, layer 1 every goal (dmfb.py:368-392): A cells each
    assert bool(((s0[:, 0] > 0).sum((1, 2)) == A).all()) and bool(((s0[:, 1] > 0).sum((1, 2)) == A).all())
    buf = P.ReplayBufferGPU(512, T, A, info["obs_shape"][-1], n_act, dev, seed=3, state_dim=3 * W * L)
    buf.store_episodes(ep)
    w0 = learner.eval_qmix_net.hyper_w1.weight.detach().clone()
    r0 = learner.eval_rnn.fc1.weight.detach().clone()
    losses = [float(learner.learn(buf.sample(32), s)) for s in range(4)]
    assert np.isfinite(losses).all()
    assert not torch.equal(w0, learner.eval_qmix_net.hyper_w1.weight) and not torch.equal(r0, learner.eval_rnn.fc1.weight)


def test_meda_training_config4_rollout_and_updates():
    """BASELINE config #4: VDN training on GPU-resident MEDA envs (30x60 chip, 4 droplets, fov 19, MEDAEnv's base
    observation - what common/config.py:10-16 hands to train.py - through the fov-19 CRNN of base_net.py:23-33)."""
    P = importlib.import_module("marl-dmfb_b200")
    dev = torch.device("cuda:0")
    N, A, n_act = 192, 4, 9
    env = P.BatchedMEDA(N, 30, 60, A, fov=19, obs_version=0, device=dev, seed=5)
    info = env.get_env_info()
    assert info["obs_shape"] == (4, 19, 19, 2, 1446) and info["episode_limit"] == 90
    T, D = info["episode_limit"], info["obs_shape"][-1]
    timer = P.PhaseTimer()
    learner = P.VDNLearner(info["obs_shape"], A, n_act, dev, hyper_hidden_dim=32, grad_norm_clip=10.0, seed=0, timer=timer)
    agents = P.BatchedAgents(learner.eval_rnn, A, n_act, dev, seed=1)
    worker = P.BatchedRolloutWorker(env, agents, epsilon=1.0, anneal_steps=300000, timer=timer)
    ep, stats = worker.generate_episodes()
    assert ep.o_all.shape == (T + 1, N, A, D)
    # the step kernel wrote straight into the episode buffer: slice t+1 is the env's observation after step t
    b = ep.as_dict()
    pad = b["padded"][:, :, 0]
    assert not bool(pad[:, 0].any()) and bool((pad[:, 1:] >= pad[:, :-1]).all())
    assert not bool(b["o_next"][pad].any()) and bool(b["o"][:, 0].any())
    buf = P.ReplayBufferGPU(256, T, A, D, n_act, dev, seed=3)
    buf.store_episodes(ep)
    w0 = learner.eval_rnn.conv2.weight.detach().clone()
    losses = [float(learner.learn(buf.sample(16), s)) for s in range(3)]
    assert np.isfinite(losses).all() and not torch.equal(w0, learner.eval_rnn.conv2.weight)
    ph = timer.summary()
    assert ph["env_step"] > 0 and ph["policy_forward"] > 0 and "grad_allreduce" in ph


def test_rollout_over_a_sub_batched_env_equals_the_single_launch_env():
    """The lock-step rollout joins after every step (the policy needs all observations), so an env stepped as K
    sub-batches on K streams (pipeline.py) must hand the worker exactly the episodes of the single-launch env."""
    P = importlib.import_module("marl-dmfb_b200")
    dev = torch.device("cuda:0")
    N, A, n_act = 600, 4, 5
    eps = []
    for K in (1, 3):
        env = P.BatchedDMFB(N, 10, 10, A, fov=9, device=dev, seed=11, sub_batches=K)
        info = env.get_env_info()
        learner = P.VDNLearner(info["obs_shape"], A, n_act, dev, seed=0)
        agents = P.BatchedAgents(learner.eval_rnn, A, n_act, dev, seed=1)
        worker = P.BatchedRolloutWorker(env, agents, epsilon=0.5, anneal_steps=1000)
        ep, stats = worker.generate_episodes()
        torch.cuda.synchronize()
        eps.append((ep, stats))
    (a, sa), (b, sb) = eps
    for name in a._FIELDS:
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
