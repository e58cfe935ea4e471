#!/usr/bin/env python
"""Generate golden traces from the UNMODIFIED reference env.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # all scenarios
    python tests/golden/make_golden.py dmfb_c1    # one scenario

Each scenario drives K independent reference chips ("envs") in lockstep for
``n_ep`` episodes of exactly ``T`` steps (the reference lets you keep stepping
after done / past max_step, dmfb.py:577-586) and records everything the hot
path consumes and produces:

 inputs   layouts[ep,K,A,4] (x,y,gx,gy read back from the reference's own task
          generator after each reset), actions[ep,T,K,A], draws[ep,T,K,A]
          (injected in place of random.random(), dmfb.py:335 / meda.py:280),
          degrade[K,W,L] (read back after construction)
 outputs  obs after reset and after every step, rewards (float64), dones,
          info.constraints, info.success, droplet positions, global state,
          usage/health matrices at episode boundaries.

The fixtures are what pins oracle/ (CPU restatement) and, through it and
directly, the CUDA path.  The files are small compressed .npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_dmfb, ref_meda = ref_shim.install()

# name -> kwargs
DMFB_SCENARIOS = {
    # C1: the north-star config (10x10, 4 droplets, fov 9)
    "dmfb_c1": dict(W=10, L=10, A=4, fov=9, stall=True, b_degrade=False, per_degrade=0.1,
                    K=12, n_ep=4, T=44, seed=101, p_goal=0.7, state_every=1),
    # C1 with degradation on a tiny chip so that health really decays within the trace
    "dmfb_c1_degrade": dict(W=10, L=10, A=4, fov=9, stall=True, b_degrade=True, per_degrade=1.0,
                            K=4, n_ep=160, T=40, seed=102, p_goal=0.5, state_every=0, obs_every=13),
    # stall=False exercises the -0.1 "new==old==0" branch (dmfb.py:345-346)
    "dmfb_c1_nostall": dict(W=10, L=10, A=4, fov=9, stall=False, b_degrade=True, per_degrade=0.5,
                            K=6, n_ep=6, T=42, seed=103, p_goal=0.8, state_every=4),
    # C2: 20x20, 10 droplets
    "dmfb_c2": dict(W=20, L=20, A=10, fov=9, stall=True, b_degrade=False, per_degrade=0.1,
                    K=6, n_ep=3, T=84, seed=104, p_goal=0.93, state_every=4),
    # C3: 50x50, 10 droplets, degrade on (evaDegre.py:37-38)
    "dmfb_c3": dict(W=50, L=50, A=10, fov=9, stall=True, b_degrade=True, per_degrade=1.0,
                    K=3, n_ep=3, T=202, seed=105, p_goal=0.93, state_every=50, obs_every=2),
    # odd shapes: non-square chip, other fovs, even fov
    "dmfb_12x15_f7": dict(W=12, L=15, A=6, fov=7, stall=True, b_degrade=True, per_degrade=0.7,
                          K=3, n_ep=70, T=56, seed=106, p_goal=0.85, state_every=28, obs_every=11),
    "dmfb_20x20_f5": dict(W=20, L=20, A=4, fov=5, stall=True, b_degrade=False, per_degrade=0.1,
                          K=4, n_ep=3, T=82, seed=107, p_goal=0.8, state_every=8),
    "dmfb_16x11_f8": dict(W=16, L=11, A=3, fov=8, stall=True, b_degrade=False, per_degrade=0.1,
                          K=4, n_ep=4, T=56, seed=108, p_goal=0.6, state_every=8),
    "dmfb_9x9_f9_a2": dict(W=9, L=9, A=2, fov=9, stall=True, b_degrade=False, per_degrade=0.1,
                           K=4, n_ep=4, T=38, seed=109, p_goal=0.6, state_every=4),
    "dmfb_30x30_f11": dict(W=30, L=30, A=7, fov=11, stall=True, b_degrade=True, per_degrade=1.0,
                           K=2, n_ep=2, T=122, seed=110, p_goal=0.8, state_every=20, obs_every=2),
    # 2x2 obstacles (dmfb.py:228-251,301-308,422-426): moves into a block are reverted, blocks appear in obs layer 2
    # at ABSOLUTE coordinates and in the global state
    "dmfb_blocks_12x12": dict(W=12, L=12, A=3, fov=9, stall=True, b_degrade=False, per_degrade=0.1, n_blocks=5,
                              K=6, n_ep=6, T=50, seed=111, p_goal=0.6, state_every=5),
    # C2 / obstacle traces in which episodes really END IN SUCCESS (the +10/+10 bonus and info['success'], dmfb.py:293-296,
    # 579-580): a constraint-avoiding policy (see _dmfb_safe_policy) instead of the mostly-greedy one
    "dmfb_c2_success": dict(W=20, L=20, A=10, fov=9, stall=True, b_degrade=False, per_degrade=0.1,
                            K=6, n_ep=3, T=84, seed=113, p_goal=1.0, state_every=12, obs_every=3, policy="safe"),
    "dmfb_blocks_success": dict(W=20, L=16, A=5, fov=5, stall=True, b_degrade=False, per_degrade=0.1, n_blocks=8,
                                K=4, n_ep=4, T=74, seed=114, p_goal=1.0, state_every=9, obs_every=3, policy="safe"),
    "dmfb_blocks_20x16_f5": dict(W=20, L=16, A=5, fov=5, stall=True, b_degrade=True, per_degrade=1.0, n_blocks=8,
                                 K=4, n_ep=4, T=74, seed=112, p_goal=0.7, state_every=9, obs_every=2),
}

MEDA_SCENARIOS = {
    "meda_c4": dict(W=30, L=60, A=4, fov=19, b_degrade=False, per_degrade=0.1,
                    K=6, n_ep=3, T=92, seed=201, p_goal=0.8, obs_every=2),
    "meda_c4_degrade": dict(W=30, L=60, A=8, fov=19, b_degrade=True, per_degrade=1.0,
                            K=3, n_ep=6, T=90, seed=202, p_goal=0.7, obs_every=6),
    "meda_80x80": dict(W=80, L=80, A=10, fov=19, b_degrade=True, per_degrade=0.6,
                       K=2, n_ep=2, T=162, seed=203, p_goal=0.85, obs_every=9),
    "meda_45x30_f9": dict(W=45, L=30, A=3, fov=9, b_degrade=False, per_degrade=0.1,
                          K=3, n_ep=3, T=76, seed=204, p_goal=0.8, obs_every=2),
    # Degradation that really happens (meda.py:302-309 getMoveProb < 1, :280 failed draws, :600-605 updateHealth):
    # the C4 chip pre-aged to usage 48 on every microelectrode (state set on the reference object after construction,
    # recorded as usage0), so that the first resets already degrade every cell a droplet has crossed three times ...
    "meda_c4_aged": dict(W=30, L=60, A=4, fov=19, b_degrade=True, per_degrade=1.0,
                         K=4, n_ep=14, T=90, seed=205, p_goal=0.75, obs_every=15, pre_usage=48),
    # ... a small chip aged the natural way, by 70 short episodes from usage 0 ...
    "meda_30x30_aged": dict(W=30, L=30, A=4, fov=19, b_degrade=True, per_degrade=0.8,
                            K=3, n_ep=70, T=60, seed=206, p_goal=0.7, obs_every=30),
    # ... and the 10-droplet set-order case on an aged 80x80 chip
    "meda_80x80_aged": dict(W=80, L=80, A=10, fov=19, b_degrade=True, per_degrade=1.0,
                            K=2, n_ep=5, T=160, seed=207, p_goal=0.8, obs_every=40, pre_usage=49),
}


def _dmfb_policy(rng, env, p_goal):
    acts = []
    for d in env.routing_manager.droplets:
        dx, dy = d.des_x - d.x, d.des_y - d.y
        cand = []
        if dx > 0:
            cand.append(1)
        if dx < 0:
            cand.append(2)
        if dy < 0:
            cand.append(3)
        if dy > 0:
            cand.append(4)
        if cand and rng.random() < p_goal:
            acts.append(int(cand[rng.integers(len(cand))]))
        else:
            acts.append(int(rng.integers(5)))
    return acts


def _dmfb_safe_policy(rng, env, p_goal):
    """Droplets are planned in index order; each takes the goal-ward move (else any move, else STALL) whose target
    cell keeps a Chebyshev distance >= 2 from every other droplet's old cell and from the cells already planned, and
    does not touch an obstacle: no static / dynamic constraint (dmfb.py:254-271) is ever incurred, so an episode that
    gets every droplet home within the step limit ends with success = 1."""
    rm = env.routing_manager
    old = [(d.x, d.y) for d in rm.droplets]
    new = list(old)
    acts = []
    delta = {0: (0, 0), 1: (1, 0), 2: (-1, 0), 3: (0, -1), 4: (0, 1)}

    def ok(i, c):
        if not (0 <= c[0] < env.width and 0 <= c[1] < env.length):
            return False
        for b in rm.blocks:
            if b.x_min <= c[0] <= b.x_max and b.y_min <= c[1] <= b.y_max:
                return False
        for j in range(len(old)):
            if j == i:
                continue
            for q in (old[j], new[j]):
                if abs(q[0] - c[0]) <= 1 and abs(q[1] - c[1]) <= 1:
                    return False
        return True

    for i, d in enumerate(rm.droplets):
        if (d.x, d.y) == (d.des_x, d.des_y):
            acts.append(int(rng.integers(5)))      # done droplets ignore their action when stall=True
            continue
        dist = lambda c: abs(c[0] - d.des_x) + abs(c[1] - d.des_y)  # noqa: E731
        order = sorted(range(1, 5), key=lambda a: (dist((d.x + delta[a][0], d.y + delta[a][1])), rng.random()))
        if rng.random() > p_goal:
            rng.shuffle(order)
        a = 0
        for cand in order:
            c = (d.x + delta[cand][0], d.y + delta[cand][1])
            if ok(i, c) and (dist(c) < dist(old[i]) or rng.random() < 0.35):
                a = cand
                break
        new[i] = (d.x + delta[a][0], d.y + delta[a][1])
        acts.append(a)
    return acts


def gen_dmfb(name, W, L, A, fov, stall, b_degrade, per_degrade, K, n_ep, T, seed, p_goal,
             state_every=1, obs_every=1, n_blocks=0, policy="greedy"):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)  # the reference draws layouts/degrade from the global numpy RNG
    import random as pyrandom
    envs, injs = [], []
    for _ in range(K):
        e = ref_dmfb.DMFBenv(W, L, A, n_blocks, fov=fov, stall=stall, b_degrade=b_degrade,
                             per_degrade=per_degrade)
        envs.append(e)
        injs.append(ref_shim.DrawInjector(ref_dmfb, e.routing_manager))
    # block retries come from python's random (dmfb.py:248-249), which every RoutingTaskManager re-seeds from the
    # clock (dmfb.py:154): seed it after the chips exist so that the fixtures are reproducible
    pyrandom.seed(seed)
    D = 3 * fov * fov + 2
    info0 = envs[0].get_env_info()
    assert info0["obs_shape"] == (3, fov, fov, 2, D)
    t_obs = list(range(0, T, obs_every))
    if (T - 1) not in t_obs:
        t_obs.append(T - 1)
    t_state = list(range(0, T, state_every)) if state_every else []

    def v1_obs(e):
        # DMFBenv_v0_1.getOneObs (dmfb.py:727-835) of the same chip: float64; the 4 layers are integral, the two
        # direction entries ((tar_y - cy) / length, (tar_x - cx) / width) are kept as float64
        o = np.stack([ref_dmfb.DMFBenv_v0_1.getOneObs(e, i) for i in range(A)])
        lay = o[:, :-2]
        assert o.dtype == np.float64 and np.all(lay == np.round(lay)) and np.abs(lay).max() < 128
        return lay.astype(np.int8), o[:, -2:].copy()

    out = dict(
        kind="dmfb", W=W, L=L, A=A, fov=fov, stall=int(stall), b_degrade=int(b_degrade),
        per_degrade=per_degrade, K=K, n_ep=n_ep, T=T, n_blocks=n_blocks,
        episode_limit=info0["episode_limit"], n_actions=info0["n_actions"],
        degrade=np.stack([e.routing_manager.m_degrade for e in envs]),
        layouts=np.zeros((n_ep, K, A, 4), np.int16),
        blocks=np.zeros((n_ep, K, n_blocks, 2), np.int16),   # (x_min, y_min) of every 2x2 block
        actions=np.zeros((n_ep, T, K, A), np.int8),
        draws=rng.random((n_ep, T, K, A)),
        draws_used=np.zeros((n_ep, T, K, A), np.uint8),
        obs_reset=np.zeros((n_ep, K, A, D), np.int8),
        obs_t=np.array(t_obs, np.int32),
        obs=np.zeros((n_ep, len(t_obs), K, A, D), np.int8),
        obs1_reset=np.zeros((n_ep, K, A, 4 * fov * fov), np.int8),
        dir1_reset=np.zeros((n_ep, K, A, 2), np.float64),
        obs1=np.zeros((n_ep, len(t_obs), K, A, 4 * fov * fov), np.int8),
        dir1=np.zeros((n_ep, len(t_obs), K, A, 2), np.float64),
        reward=np.zeros((n_ep, T, K, A), np.float64),
        done=np.zeros((n_ep, T, K, A), np.uint8),
        constraints=np.zeros((n_ep, T, K), np.int32),
        success=np.zeros((n_ep, T, K), np.uint8),
        pos=np.zeros((n_ep, T, K, A, 2), np.int16),
        state_t=np.array(t_state, np.int32),
        state=np.zeros((n_ep, len(t_state), K, 3, W, L), np.int8),
        health_reset=np.zeros((n_ep, K, W, L), np.float64),
        usage_reset=np.zeros((n_ep, K, W, L), np.float64),
        usage_end=np.zeros((n_ep, K, W, L), np.float64),
    )
    for ep in range(n_ep):
        for k, e in enumerate(envs):
            obs = e.reset()  # new=False: keeps health, applies updateHealth (dmfb.py:589-597)
            rm = e.routing_manager
            out["layouts"][ep, k, :, 0:2] = rm.starts
            out["layouts"][ep, k, :, 2:4] = rm.ends
            assert len(rm.blocks) == n_blocks
            for b, blk in enumerate(rm.blocks):
                assert (blk.x_max, blk.y_max) == (blk.x_min + 1, blk.y_min + 1)
                out["blocks"][ep, k, b] = (blk.x_min, blk.y_min)
            out["obs_reset"][ep, k] = np.stack(obs)
            assert all(o.dtype == np.int8 for o in obs)
            out["obs1_reset"][ep, k], out["dir1_reset"][ep, k] = v1_obs(e)
            out["health_reset"][ep, k] = rm.m_health
            out["usage_reset"][ep, k] = rm.m_usage
        for t in range(T):
            for k, e in enumerate(envs):
                acts = (_dmfb_safe_policy if policy == "safe" else _dmfb_policy)(rng, e, p_goal)
                out["actions"][ep, t, k] = acts
                obs, rew, done, info = injs[k].step(e, acts, out["draws"][ep, t, k])
                out["draws_used"][ep, t, k] = injs[k].consumed
                if t in t_obs:
                    out["obs"][ep, t_obs.index(t), k] = np.stack(obs)
                    out["obs1"][ep, t_obs.index(t), k], out["dir1"][ep, t_obs.index(t), k] = v1_obs(e)
                out["reward"][ep, t, k] = [rew[a] for a in e.agents]
                out["done"][ep, t, k] = [done[a] for a in e.agents]
                out["constraints"][ep, t, k] = info["constraints"]
                out["success"][ep, t, k] = info["success"]
                out["pos"][ep, t, k] = [(d.x, d.y) for d in e.routing_manager.droplets]
                if t in t_state:
                    g = e.routing_manager.global_obs
                    assert g.min() >= 0 and g.max() <= 127
                    out["state"][ep, t_state.index(t), k] = g
        for k, e in enumerate(envs):
            out["usage_end"][ep, k] = e.routing_manager.m_usage
    out["health_final"] = np.stack([e.routing_manager.m_health for e in envs])
    return out


def _meda_policy(rng, env, p_goal):
    acts = []
    rm = env.routing_manager
    for d, g in zip(rm.droplets, rm.destinations):
        dx, dy = g.x_center - d.x_center, g.y_center - d.y_center
        if rng.random() < p_goal and (dx or dy):
            if abs(dx) >= 2 and abs(dy) >= 2:
                a = {(1, -1): 4, (1, 1): 5, (-1, 1): 6, (-1, -1): 7}[(int(np.sign(dx)), int(np.sign(dy)))]
            elif abs(dx) >= abs(dy):
                a = 1 if dx > 0 else 3
            else:
                a = 2 if dy > 0 else 0
            acts.append(a)
        else:
            acts.append(int(rng.integers(9)))
    return acts


def gen_meda(name, W, L, A, fov, b_degrade, per_degrade, K, n_ep, T, seed, p_goal, obs_every=1, pre_usage=0):
    import random as pyrandom
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    envs, injs = [], []
    for k in range(K):
        # MEDAEnv_v0_2 only overrides getOneObs (meda.py:846-897); the base-class obs is
        # taken from the same object through MEDAEnv.getOneObs (meda.py:613-674).
        e = ref_meda.MEDAEnv_v0_2(W, L, A, fov=fov, b_degrade=b_degrade, per_degrade=per_degrade)
        pyrandom.seed(seed * 1000 + k)  # task generator uses python's random (meda.py:224-227)
        if pre_usage:
            e.m_usage[...] = float(pre_usage)   # aged chip: m_usage is plain env state (meda.py:495)
        envs.append(e)
        injs.append(ref_shim.DrawInjector(ref_meda, e.routing_manager))
    # bookkeeping for the printout only: the probabilities getMoveProb returned (instance wrapper, class untouched)
    probs_seen = []
    for e in envs:
        def _wrap(orig):
            def f(droplet, m_health):
                p = orig(droplet, m_health)
                probs_seen.append(p)
                return p
            return f
        e.routing_manager.getMoveProb = _wrap(e.routing_manager.getMoveProb)
    D0 = 4 * fov * fov + 2
    D2 = 3 * fov * fov + 2
    t_obs = list(range(0, T, obs_every))
    if (T - 1) not in t_obs:
        t_obs.append(T - 1)

    def base_obs(e):
        o = np.stack([ref_meda.MEDAEnv.getOneObs(e, i) for i in range(A)])
        assert o.dtype == np.float64 and np.all(o == np.round(o)) and np.abs(o).max() < 128
        return o.astype(np.int8)

    def v2_obs(e):
        o = np.stack([e.getOneObs(i) for i in range(A)])
        assert o.dtype == np.int8
        return o

    def v1_obs(e):
        # MEDAEnv_v0_1.getOneObs (meda.py:788-844): float64; the 4 layers are integral, the two direction
        # entries are (dy / width, dx / length) and are kept as float64
        o = np.stack([ref_meda.MEDAEnv_v0_1.getOneObs(e, i) for i in range(A)])
        lay = o[:, :-2]
        assert o.dtype == np.float64 and np.all(lay == np.round(lay)) and np.abs(lay).max() < 128
        return lay.astype(np.int8), o[:, -2:].copy()

    out = dict(
        kind="meda", W=W, L=L, A=A, fov=fov, b_degrade=int(b_degrade), per_degrade=per_degrade,
        K=K, n_ep=n_ep, T=T, episode_limit=envs[0].max_step, n_actions=9,
        degrade=np.stack([e.m_degrade for e in envs]),
        usage0=np.stack([e.m_usage for e in envs]),   # state before the first reset (nonzero for pre-aged chips)
        layouts=np.zeros((n_ep, K, A, 4), np.int16),  # x_c, y_c, goal x_c, goal y_c
        actions=np.zeros((n_ep, T, K, A), np.int8),
        draws=rng.random((n_ep, T, K, A)),
        draws_used=np.zeros((n_ep, T, K, A), np.uint8),
        obs_t=np.array(t_obs, np.int32),
        obs0_reset=np.zeros((n_ep, K, A, D0), np.int8),
        obs2_reset=np.zeros((n_ep, K, A, D2), np.int8),
        obs0=np.zeros((n_ep, len(t_obs), K, A, D0), np.int8),
        obs2=np.zeros((n_ep, len(t_obs), K, A, D2), np.int8),
        obs1_reset=np.zeros((n_ep, K, A, D0 - 2), np.int8),
        dir1_reset=np.zeros((n_ep, K, A, 2), np.float64),
        obs1=np.zeros((n_ep, len(t_obs), K, A, D0 - 2), np.int8),
        dir1=np.zeros((n_ep, len(t_obs), K, A, 2), np.float64),
        reward=np.zeros((n_ep, T, K, A), np.float64),
        done=np.zeros((n_ep, T, K, A), np.uint8),
        status=np.zeros((n_ep, T, K, A), np.uint8),
        constraints=np.zeros((n_ep, T, K), np.float64),
        success=np.zeros((n_ep, T, K), np.uint8),
        pos=np.zeros((n_ep, T, K, A, 2), np.int16),
        health_reset=np.zeros((n_ep, K, W, L), np.float64),
        usage_reset=np.zeros((n_ep, K, W, L), np.float64),
        usage_end=np.zeros((n_ep, K, W, L), np.float64),
    )
    for ep in range(n_ep):
        for k, e in enumerate(envs):
            obs = e.reset()
            rm = e.routing_manager
            out["layouts"][ep, k] = [(d.x_center, d.y_center, g.x_center, g.y_center)
                                     for d, g in zip(rm.droplets, rm.destinations)]
            out["obs2_reset"][ep, k] = np.stack(obs)
            out["obs0_reset"][ep, k] = base_obs(e)
            out["obs1_reset"][ep, k], out["dir1_reset"][ep, k] = v1_obs(e)
            out["health_reset"][ep, k] = e.m_health
            out["usage_reset"][ep, k] = e.m_usage
        for t in range(T):
            for k, e in enumerate(envs):
                acts = _meda_policy(rng, e, p_goal)
                out["actions"][ep, t, k] = acts
                obs, rew, done, info = injs[k].step(e, acts, out["draws"][ep, t, k])
                out["draws_used"][ep, t, k] = injs[k].consumed
                if t in t_obs:
                    out["obs2"][ep, t_obs.index(t), k] = np.stack(obs)
                    out["obs0"][ep, t_obs.index(t), k] = base_obs(e)
                    out["obs1"][ep, t_obs.index(t), k], out["dir1"][ep, t_obs.index(t), k] = v1_obs(e)
                out["reward"][ep, t, k] = [rew[a] for a in e.agents]
                out["done"][ep, t, k] = [done[a] for a in e.agents]
                out["status"][ep, t, k] = e.routing_manager.status
                out["constraints"][ep, t, k] = info["constraints"]
                out["success"][ep, t, k] = info["success"]
                out["pos"][ep, t, k] = [(d.x_center, d.y_center) for d in e.routing_manager.droplets]
        for k, e in enumerate(envs):
            out["usage_end"][ep, k] = e.m_usage
    out["health_final"] = np.stack([e.m_health for e in envs])
    # draws are consumed in the order getMoveProb was called: one per True entry of draws_used, (ep, t, k, i) order
    used = out["draws"][out["draws_used"] > 0]
    assert len(used) == len(probs_seen)
    out["_failed_draws"] = int((used > np.array(probs_seen)).sum())
    out["_min_prob"] = float(min(probs_seen)) if probs_seen else 1.0
    return out


def main(argv):
    want = set(argv[1:])
    for name, kw in DMFB_SCENARIOS.items():
        if want and name not in want:
            continue
        data = gen_dmfb(name, **kw)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  success-steps={int(data['success'].sum())} "
              f"constraint-steps={int((data['constraints'] > 0).sum())} "
              f"min-health={data['health_final'].min():.3g}")
    for name, kw in MEDA_SCENARIOS.items():
        if want and name not in want:
            continue
        data = gen_meda(name, **kw)
        failed, min_prob = data.pop("_failed_draws"), data.pop("_min_prob")
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  success-steps={int(data['success'].sum())} "
              f"punish-steps={int((data['constraints'] < 0).sum())} "
              f"min-health={data['health_final'].min():.3g} degraded-cells={int((data['health_final'] < 1).sum())} "
              f"failed-draws={failed} min-move-prob={min_prob:.3g}")


if __name__ == "__main__":
    main(sys.argv)
