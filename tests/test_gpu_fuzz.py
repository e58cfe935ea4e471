"""Randomised parity sweep: random chip sizes / droplet counts / fov / obstacle counts / options, CUDA env vs the CPU
oracle, bit for bit, two episodes each (DMFB base + v0_1 observations, MEDA base / v0_1 / v0_2)."""
import importlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

npy = lambda t: t.detach().cpu().numpy()  # noqa: E731


def eq(a, b, what, cfg):
    assert np.array_equal(a, b), f"MISMATCH {what} for {cfg}"


def dmfb_case(P, O, rng):
    W, L = int(rng.integers(5, 41)), int(rng.integers(5, 41))
    fov = int(rng.integers(1, min(W, L, 19) + 1))
    if fov // 2 == 10:
        fov -= 2
    # droplet density bounded so that the whole-set rejection sampler of the oracle accepts within ~e^4 attempts
    A = int(rng.integers(2, max(2, min(32, int(0.6 * np.sqrt(W * L)))) + 1))
    # obstacles only where they can always be placed (GenRandomBlocks retries for ever otherwise, dmfb.py:246-250)
    nb = int(rng.integers(0, 4)) * int(rng.integers(0, 2)) if min(W, L) >= 12 else 0
    stall, deg, ver = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), int(rng.integers(0, 2))
    N = int(rng.integers(1, 200))
    cfg = dict(kind="dmfb", N=N, W=W, L=L, A=A, fov=fov, nb=nb, stall=stall, deg=deg, ver=ver)
    ref = O.OracleDMFB(N, W, L, A, fov=fov, stall=stall, b_degrade=deg, n_blocks=nb, obs_version=ver)
    try:
        lay = ref.gen_layouts(seed=int(rng.integers(1 << 30)))
    except Exception:
        return None
    blocks = ref.gen_blocks(int(rng.integers(1 << 30)), lay) if nb else None
    degrade = rng.random((N, W, L)) * 0.4 + 0.6 if deg else None
    env = P.BatchedDMFB(N, W, L, A, nb, fov=fov, stall=stall, b_degrade=deg, per_degrade=1.0, device="cuda:0", track_usage=True,
                        reward_f64=True, degrade=degrade, layouts=lay, block_layouts=blocks, obs_version=ver)
    if deg:
        ref.degrade[...] = degrade
        ref.health[...] = rng.random((N, W, L)) * 0.7 + 0.3
        env.health.copy_(torch.as_tensor(ref.health))
        ref.usage[...] = 45.0
        env.usage.fill_(45)
    for ep in range(2):
        lay = ref.gen_layouts(seed=int(rng.integers(1 << 30)))
        blocks = ref.gen_blocks(int(rng.integers(1 << 30)), lay) if nb else None
        eq(npy(env.reset(layouts=lay, block_layouts=blocks)), ref.reset(lay, blocks=blocks), "reset obs", cfg)
        for t in range(int(rng.integers(3, 2 * (W + L) + 6))):
            d = ref.drop.astype(np.int32)
            dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
            toward = np.where(np.abs(dx) >= np.abs(dy), np.where(dx > 0, 1, 2), np.where(dy > 0, 4, 3))
            acts = np.where(rng.random((N, A)) < 0.6, toward, rng.integers(0, 5, (N, A))).astype(np.int8)
            draws = rng.random((N, A))
            obs, rew, done, cons, succ = ref.step(acts, draws)
            g_obs, _, g_done, info = env.step(torch.as_tensor(acts, device="cuda:0"), draws=draws)
            eq(npy(env.drop), ref.drop, f"drop t{t}", cfg)
            eq(npy(g_obs), obs, f"obs t{t}", cfg)
            eq(npy(env.reward_f64), rew, f"reward t{t}", cfg)
            eq(npy(g_done).astype(np.uint8), done, f"done t{t}", cfg)
            eq(npy(info["constraints"]), cons, f"constraints t{t}", cfg)
            eq(npy(info["success"]), succ, f"success t{t}", cfg)
        eq(npy(env.usage_counts()), ref.usage, "usage", cfg)
        eq(npy(env.get_state()), ref.global_state(), "state", cfg)
        if deg:
            eq(npy(env.health), ref.health, "health", cfg)
    return cfg


def meda_case(P, O, rng):
    W, L = int(rng.integers(15, 91)), int(rng.integers(15, 91))
    limit = (W // 15) * (L // 15)
    A = int(rng.integers(1, min(limit, 16) + 1))
    fov = int(rng.choice([5, 7, 9, 11, 13, 15, 19, 21]))
    deg, ver = bool(rng.integers(0, 2)), int(rng.integers(0, 3))
    N = int(rng.integers(1, 120))
    cfg = dict(kind="meda", N=N, W=W, L=L, A=A, fov=fov, deg=deg, ver=ver)
    ref = O.OracleMEDA(N, W, L, A, fov=fov, b_degrade=deg, obs_version=ver)
    try:
        lay = ref.gen_layouts(seed=int(rng.integers(1 << 30)))
    except Exception:
        return None
    degrade = rng.random((N, W, L)) * 0.4 + 0.6 if deg else None
    env = P.BatchedMEDA(N, W, L, A, fov=fov, b_degrade=deg, per_degrade=1.0, obs_version=ver, device="cuda:0", reward_f64=True,
                        degrade=degrade, layouts=lay, track_usage=True)
    if deg:
        ref.degrade[...] = degrade
        ref.health[...] = rng.random((N, W, L)) * 0.7 + 0.3
        env.health.copy_(torch.as_tensor(ref.health))
        ref.usage[...] = 44.0
        env.usage.fill_(44)
    for ep in range(2):
        lay = ref.gen_layouts(seed=int(rng.integers(1 << 30)))
        eq(npy(env.reset(layouts=lay)), ref.reset(lay), "reset obs", cfg)
        if deg:
            eq(npy(env.health), ref.health, "health after reset", cfg)
        for t in range(int(rng.integers(3, W + L + 6))):
            d = ref.drop.astype(np.int32)
            dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
            toward = np.where(np.abs(dx) >= np.abs(dy), np.where(dx > 0, 1, 3), np.where(dy > 0, 2, 0))
            acts = np.where(rng.random((N, A)) < 0.7, toward, rng.integers(0, 9, (N, A))).astype(np.int8)
            draws = rng.random((N, A))
            obs, rew, done, cons, succ = ref.step(acts, draws)
            g_obs, _, g_done, info = env.step(torch.as_tensor(acts, device="cuda:0"), draws=draws)
            eq(npy(env.drop), ref.drop, f"drop t{t}", cfg)
            eq(npy(g_obs), obs, f"obs t{t}", cfg)
            eq(npy(env.reward_f64), rew, f"reward t{t}", cfg)
            eq(npy(g_done).astype(np.uint8), done, f"done t{t}", cfg)
            eq(npy(info["constraints"]), cons, f"punish t{t}", cfg)
            eq(npy(info["success"]), succ, f"success t{t}", cfg)
        eq(npy(env.usage_counts()), ref.usage, "usage", cfg)
    return cfg


@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_configurations_match_the_oracle(oracle_lib, seed):
    P = importlib.import_module("marl-dmfb_b200")
    O = oracle_lib
    rng = np.random.default_rng(seed)
    done = []
    while len(done) < 10:
        c = dmfb_case(P, O, rng) if rng.random() < 0.55 else meda_case(P, O, rng)
        if c is not None:
            done.append(c)
    assert {c["kind"] for c in done} == {"dmfb", "meda"}
