"""Host restatement of the library's device task generator (marl-dmfb_b200/csrc/dmfb_kernels.cu: layout_stream,
sample_rounds_warp / search_env_lanewise for 10 droplets, sample_round otherwise; common.cuh: mix64) in numpy.

_Generate_Start_End (dmfb.py:207-226) draws 2A uniform cells and redraws the whole set until no two of them are equal or
8-adjacent.  On the device attempt k of (seed, env, episode) is a pure function of those values and the task is the
FIRST accepted attempt; this file computes exactly that, so that the tasks the kernels produce - whichever kernel
or flavour found them, in whichever step - can be compared bit for bit."""
import numpy as np

PHI = np.uint64(0x9E3779B97F4A7C15)
C_ENV, C_EPI = np.uint64(0xD1342543DE82EF95), np.uint64(0xDA942042E4DD58B5)
STREAM_LAYOUT = 2
U32 = np.uint64(0xFFFFFFFF)


def mix64(z):
    z = np.asarray(z, np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _cells(z, W, L):
    """x = umulhi(low word, W), y = umulhi(high word, L)"""
    return ((z & U32) * np.uint64(W)) >> np.uint64(32), ((z >> np.uint64(32)) * np.uint64(L)) >> np.uint64(32)


def _accepted(x, y):
    """[n, P] points -> [n] bool: every pair differs by more than one cell in x or in y"""
    dx = np.abs(x[:, :, None].astype(np.int64) - x[:, None, :].astype(np.int64))
    dy = np.abs(y[:, :, None].astype(np.int64) - y[:, None, :].astype(np.int64))
    near = (dx <= 1) & (dy <= 1)
    near[:, np.arange(x.shape[1]), np.arange(x.shape[1])] = False
    return ~near.any(axis=(1, 2))


def first_accepted_tasks(seed, env, episode, W, L, A, max_attempts=100000):
    """uint8 [N, A, 4] = (x, y, goal_x, goal_y) and the attempt index [N] of the first accepted attempt."""
    with np.errstate(over="ignore"):
        env = np.asarray(env, np.uint64)
        epi = np.asarray(episode, np.uint64)
        n = env.shape[0]
        base0 = np.uint64(seed) ^ (PHI * np.uint64(STREAM_LAYOUT + 1))
        out = np.zeros((n, A, 4), np.uint8)
        at = np.full(n, -1, np.int64)
        todo = np.arange(n)
        if A == 10:     # one stream per (env, episode); point p of attempt k: counter k * 2A + p + 1
            base = mix64((base0 + env * C_ENV + (epi << np.uint64(32)) * C_EPI) ^ np.uint64(0xA5A5A5A5A5A5A5A5))
            p1 = (np.arange(2 * A, dtype=np.uint64) + np.uint64(1))[None, :] * PHI
        else:           # one stream per (env, episode, droplet); attempt k: counters 2k + 1 (start), 2k + 2 (goal)
            i = np.arange(A, dtype=np.uint64)[None, :]
            base = mix64(base0 + env[:, None] * C_ENV + ((epi[:, None] << np.uint64(32)) | i) * C_EPI)
        for k in range(max_attempts):
            if todo.size == 0:
                break
            kk = np.uint64(k & 0xFFFFFFFF)
            if A == 10:
                z = mix64((base[todo] + kk * np.uint64(2 * A) * PHI)[:, None] + p1)
                x, y = _cells(z, W, L)                                   # [n, 2A]: point 2i = start, 2i + 1 = goal
            else:
                z0 = mix64(base[todo] + (np.uint64(2) * kk + np.uint64(1)) * PHI)
                z1 = mix64(base[todo] + (np.uint64(2) * kk + np.uint64(2)) * PHI)
                (sx, sy), (gx, gy) = _cells(z0, W, L), _cells(z1, W, L)
                x, y = np.stack([sx, gx], 2).reshape(todo.size, 2 * A), np.stack([sy, gy], 2).reshape(todo.size, 2 * A)
            ok = _accepted(x, y)
            sel = todo[ok]
            out[sel, :, 0], out[sel, :, 1] = x[ok][:, 0::2], y[ok][:, 0::2]
            out[sel, :, 2], out[sel, :, 3] = x[ok][:, 1::2], y[ok][:, 1::2]
            at[sel] = k
            todo = todo[~ok]
        assert todo.size == 0, "density that cannot be placed within max_attempts"
        return out, at


def meda_first_tasks(seed, env, episode, W, L, A, r=2, max_draws=100000):
    """meda_generate_tasks (marl-dmfb_b200/csrc/meda_kernels.cu), the device form of refresh / addTask /
    _genLegalDroplet (meda.py:161-185,213-233): ONE sequential splitmix64 stream per (seed, env, episode); every draw
    is a centre (y from the low word against `width`, x from the high word against `length`, both in [r, dim-r-1]);
    droplet i's start is redrawn while it is closer than 9 to an earlier start, its destination while it is closer
    than 9 to an earlier destination or overlaps its own start.  uint8 [N, A, 4] = (x, y, goal_x, goal_y)."""
    out = np.zeros((len(env), A, 4), np.uint8)
    with np.errstate(over="ignore"):
        for n, (e, ep) in enumerate(zip(np.asarray(env, np.uint64), np.asarray(episode, np.uint64))):
            state = mix64((np.uint64(seed) ^ (PHI * np.uint64(STREAM_LAYOUT + 1))) + e * C_ENV + (ep << np.uint64(32)) * C_EPI)
            draws = 0

            def centre():
                nonlocal state, draws
                draws += 1
                assert draws < max_draws
                state = state + PHI
                z = mix64(state)
                y = r + ((int(z) & 0xFFFFFFFF) * (W - 2 * r) >> 32)
                x = r + ((int(z) >> 32) * (L - 2 * r) >> 32)
                return x, y

            for i in range(A):
                while True:
                    sx, sy = centre()
                    if all((sx - int(out[n, j, 0])) ** 2 + (sy - int(out[n, j, 1])) ** 2 >= 81 for j in range(i)):
                        break
                while True:
                    tx, ty = centre()
                    if any((tx - int(out[n, j, 2])) ** 2 + (ty - int(out[n, j, 3])) ** 2 < 81 for j in range(i)):
                        continue
                    if abs(tx - sx) <= 2 * r and abs(ty - sy) <= 2 * r:      # isDropletOverlap (meda.py:71-81)
                        continue
                    break
                out[n, i] = (sx, sy, tx, ty)
    return out


STREAM_BLOCKS = 4


def first_blocks(seed, env, episode, tasks, W, L, n_blocks, max_draws=100000):
    """generate_blocks (marl-dmfb_b200/csrc/dmfb_kernels.cu), the device form of GenRandomBlocks (dmfb.py:228-251): one
    sequential splitmix64 stream per (seed, env, episode); every draw is the (x_min, y_min) of a 2x2 block, uniform in
    [0, W-4] x [0, L-4]; it is redrawn while the block covers a start or goal cell of the env's task `tasks[n]`
    (uint8 [A, 4] = x, y, goal_x, goal_y) or overlaps / touches an earlier block (isBlockOverlap, dmfb.py:56-69).
    uint8 [N, n_blocks, 2]."""
    out = np.zeros((len(env), n_blocks, 2), np.uint8)
    with np.errstate(over="ignore"):
        for n, (e, ep) in enumerate(zip(np.asarray(env, np.uint64), np.asarray(episode, np.uint64))):
            state = mix64((np.uint64(seed) ^ (PHI * np.uint64(STREAM_BLOCKS + 1))) + e * C_ENV + (ep << np.uint64(32)) * C_EPI)
            cells = [(int(t[0]), int(t[1])) for t in tasks[n]] + [(int(t[2]), int(t[3])) for t in tasks[n]]
            draws = 0
            for b in range(n_blocks):
                while True:
                    draws += 1
                    assert draws < max_draws
                    state = state + PHI
                    z = int(mix64(state))
                    x, y = ((z & 0xFFFFFFFF) * (W - 3)) >> 32, ((z >> 32) * (L - 3)) >> 32
                    if any(0 <= cx - x <= 1 and 0 <= cy - y <= 1 for cx, cy in cells):
                        continue
                    if any(not (x > int(out[n, k, 0]) + 1 or int(out[n, k, 0]) > x + 1) and
                           not (y > int(out[n, k, 1]) + 1 or int(out[n, k, 1]) > y + 1) for k in range(b)):
                        continue
                    out[n, b] = (x, y)
                    break
    return out
