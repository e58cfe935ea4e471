"""Host restatement of the library's counter-based move-success draws (marl-dmfb_b200/csrc/common.cuh:
philox4x32_10, env_random, u53) in numpy, so that the device-RNG path can be pinned against injected draws."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_MOVE = 1


def philox4x32_10(ctr, key):
    """ctr: 4 uint64 arrays holding 32-bit values, key: 2 python ints / arrays.  Salmon et al. 2011."""
    x, y, z, w = [np.asarray(c, np.uint64) & MASK for c in ctr]
    k0 = np.asarray(key[0], np.uint64) & MASK
    k1 = np.asarray(key[1], np.uint64) & MASK
    for _ in range(10):
        p0, p1 = M0 * x, M1 * z
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        x, y, z, w = (hi1 ^ y ^ k0) & MASK, lo1, (hi0 ^ w ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return x, y, z, w


def move_draws(seed, env, episode, step, n_agents):
    """float64 [N, A]: the draw of droplet i of global env `env[n]` in its `episode[n]` at step `step[n]` (1-based)."""
    env = np.asarray(env, np.uint64)[:, None]
    episode = np.asarray(episode, np.uint64)[:, None]
    step = np.asarray(step, np.uint64)[:, None]
    agent = np.arange(n_agents, dtype=np.uint64)[None, :]
    key0 = (seed & 0xFFFFFFFF) ^ ((STREAM_MOVE * 0x85EBCA6B) & 0xFFFFFFFF)
    key1 = (np.uint64((seed >> 32) & 0xFFFFFFFF) ^ (env >> np.uint64(32))) & MASK
    z = np.zeros_like(env + agent)
    x, y, _, _ = philox4x32_10(((env & MASK) + z, episode + z, step + z, agent + z), (np.uint64(key0) + z, key1 + z))
    return ((x >> np.uint64(5)).astype(np.float64) * 67108864.0 + (y >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


STREAM_DEGRADE = 3


def degrade_matrix(seed, env, episode, cells, per_degrade):
    """float64 [N, cells]: the device form of _random_health_statue (dmfb.py:157-164, meda.py:494-504) - cell k of global
    env `env[n]` draws Philox(counter = env, episode, k, 0; stream 3): degrade = u53(x, y) * 0.4 + 0.6 (two roundings, as
    NumPy computes rand * 0.4 + 0.6), set to 1.0 where u53(z, w) < 1 - per_degrade."""
    env = np.asarray(env, np.uint64)[:, None]
    episode = np.asarray(episode, np.uint64)[:, None]
    k = np.arange(cells, dtype=np.uint64)[None, :]
    key0 = (seed & 0xFFFFFFFF) ^ ((STREAM_DEGRADE * 0x85EBCA6B) & 0xFFFFFFFF)
    key1 = (np.uint64((seed >> 32) & 0xFFFFFFFF) ^ (env >> np.uint64(32))) & MASK
    z = np.zeros_like(env + k)
    x, y, zz, w = philox4x32_10(((env & MASK) + z, episode + z, k + z, z), (np.uint64(key0) + z, key1 + z))
    u53 = lambda a, b: ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0  # noqa: E731
    dg = u53(x, y) * 0.4 + 0.6
    return np.where(u53(zz, w) < 1.0 - per_degrade, 1.0, dg)
