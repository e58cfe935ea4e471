"""GPU parity for the MEDA hot path: CUDA kernels (through the C ABI) vs golden traces recorded from the
unmodified reference env and vs the CPU oracle on seeded random inputs.  Bit-exact for positions, status,
dones, observations (both MEDAEnv and MEDAEnv_v0_2 variants), usage, health and the float64 rewards."""
import importlib

import numpy as np
import pytest

from conftest import golden_names, load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def pkg():
    return importlib.import_module("marl-dmfb_b200")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", golden_names("meda"))
def test_meda_cuda_matches_reference_trace(name):
    g = load_golden(name)
    K, A, W, L = g["K"], g["A"], g["W"], g["L"]
    env = pkg().BatchedMEDA(K, W, L, A, fov=g["fov"], b_degrade=bool(g["b_degrade"]), per_degrade=g["per_degrade"],
                            obs_version=2, device="cuda:0", reward_f64=True, track_usage=True,
                            degrade=g["degrade"] if g["b_degrade"] else None, layouts=g["layouts"][0])
    base = pkg().BatchedMEDA(K, W, L, A, fov=g["fov"], obs_version=0, device="cuda:0", layouts=g["layouts"][0])
    v01 = pkg().BatchedMEDA(K, W, L, A, fov=g["fov"], obs_version=1, device="cuda:0", layouts=g["layouts"][0])

    def check_v01(layers, dirs, msg):
        v01.drop.copy_(env.drop)
        o = _np(v01.get_obs())
        np.testing.assert_array_equal(o[..., :-2], layers, err_msg=msg + " obs v0_1 layers")
        np.testing.assert_array_equal(o[..., -2] / W, dirs[..., 0], err_msg=msg + " obs v0_1 dir y")
        np.testing.assert_array_equal(o[..., -1] / L, dirs[..., 1], err_msg=msg + " obs v0_1 dir x")

    obs_t = list(g["obs_t"])
    env.usage.copy_(torch.as_tensor(g["usage0"].astype(np.int32)))   # pre-aged chips (meda_*_aged)
    for ep in range(g["n_ep"]):
        obs = env.reset(layouts=g["layouts"][ep])
        np.testing.assert_array_equal(_np(obs), g["obs2_reset"][ep], err_msg=f"{name} v0_2 reset obs ep{ep}")
        base.drop.copy_(env.drop)
        np.testing.assert_array_equal(_np(base.get_obs()), g["obs0_reset"][ep], err_msg=f"{name} base reset obs ep{ep}")
        check_v01(g["obs1_reset"][ep], g["dir1_reset"][ep], f"{name} reset ep{ep}")
        if g["b_degrade"]:
            np.testing.assert_array_equal(_np(env.health), g["health_reset"][ep], err_msg=f"health ep{ep}")
        np.testing.assert_array_equal(_np(env.usage_counts()), g["usage_reset"][ep], err_msg=f"usage ep{ep}")
        for t in range(g["T"]):
            acts = torch.as_tensor(g["actions"][ep, t], device="cuda:0")
            obs, rew, done, info = env.step(acts, draws=g["draws"][ep, t])
            msg = f"{name} ep{ep} t{t}"
            np.testing.assert_array_equal(_np(env.drop[:, :, 0:2]), g["pos"][ep, t], err_msg=msg + " pos")
            np.testing.assert_array_equal(_np(env.reward_f64), g["reward"][ep, t], err_msg=msg + " reward f64")
            np.testing.assert_allclose(_np(rew), g["reward"][ep, t], rtol=1e-6, atol=0, err_msg=msg + " reward f32")
            np.testing.assert_array_equal(_np(done).astype(np.uint8), g["done"][ep, t], err_msg=msg + " done")
            np.testing.assert_array_equal(_np(env.status), g["status"][ep, t], err_msg=msg + " status")
            np.testing.assert_allclose(-0.6 * _np(info["constraints"]), g["constraints"][ep, t], rtol=1e-12, atol=1e-12)
            np.testing.assert_array_equal(_np(info["success"]), g["success"][ep, t], err_msg=msg + " success")
            if t in obs_t:
                np.testing.assert_array_equal(_np(obs), g["obs2"][ep, obs_t.index(t)], err_msg=msg + " obs v0_2")
                base.drop.copy_(env.drop)
                np.testing.assert_array_equal(_np(base.get_obs()), g["obs0"][ep, obs_t.index(t)], err_msg=msg + " obs base")
                check_v01(g["obs1"][ep, obs_t.index(t)], g["dir1"][ep, obs_t.index(t)], msg)
        np.testing.assert_array_equal(_np(env.usage_counts()), g["usage_end"][ep], err_msg=f"usage end ep{ep}")
    if g["b_degrade"]:
        np.testing.assert_array_equal(_np(env.health), g["health_final"])


CASES = [
    # N, W, L, A, fov, degrade, obs_version
    (300, 30, 60, 4, 19, True, 0),
    (300, 30, 60, 4, 19, True, 2),
    (65, 30, 60, 8, 19, False, 2),
    (33, 80, 80, 10, 19, True, 2),     # A > 8: python-set write order matters
    (33, 80, 80, 10, 19, False, 0),
    (70, 45, 30, 3, 9, True, 2),
    (50, 60, 45, 6, 13, True, 0),
    (1, 30, 60, 4, 19, False, 2),
    (9, 128, 128, 32, 19, True, 0),    # the library's maxima: chip side, droplets (one env per warp)
    (7, 120, 120, 16, 19, False, 2),   # largest python-set order table (2^16 rows)
    (20, 90, 90, 15, 19, True, 0),     # long observation rows: the reset tile is capped by shared memory
    (20, 75, 90, 15, 21, False, 1),
    (300, 30, 60, 4, 19, False, 1),    # MEDAEnv_v0_1 through the step kernel (specialised instance)
    (33, 80, 80, 10, 19, True, 1),     # ... and the generic instance with the python-set order table
    (70, 45, 30, 3, 9, False, 1),
]


@pytest.mark.parametrize("N,W,L,A,fov,deg,ver", CASES)
def test_meda_cuda_matches_oracle_random(oracle_lib, N, W, L, A, fov, deg, ver):
    rng = np.random.default_rng(N + 13 * A + ver)
    ref = oracle_lib.OracleMEDA(N, W, L, A, fov=fov, b_degrade=deg, obs_version=ver)
    degrade = rng.random((N, W, L)) * 0.4 + 0.6 if deg else None
    layouts = ref.gen_layouts(seed=N)
    env = pkg().BatchedMEDA(N, W, L, A, fov=fov, b_degrade=deg, per_degrade=1.0, obs_version=ver, device="cuda:0",
                            reward_f64=True, degrade=degrade, layouts=layouts, track_usage=True)
    if deg:
        ref.degrade[...] = degrade
        ref.health[...] = rng.random((N, W, L)) * 0.7 + 0.3
        env.health.copy_(torch.as_tensor(ref.health))
        ref.usage[...] = 40.0
        env.usage.fill_(40)
    T = min(W + L + 3, 60)
    for ep in range(3):
        layouts = ref.gen_layouts(seed=77 * ep + N)
        mask = (np.arange(N) % 3 != 0).astype(np.uint8) if ep == 1 else None
        o_ref = ref.reset(layouts, mask=mask)
        buf = env.obs.clone()
        o_gpu = env.reset(layouts=layouts, mask=mask)
        sel = slice(None) if mask is None else mask.astype(bool)
        np.testing.assert_array_equal(_np(o_gpu)[sel], o_ref[sel], err_msg=f"reset obs ep{ep}")
        if mask is not None:
            np.testing.assert_array_equal(_np(o_gpu)[~sel], _np(buf)[~sel])
        np.testing.assert_array_equal(_np(env.drop), ref.drop)
        if deg:
            np.testing.assert_array_equal(_np(env.health), ref.health)
        for t in range(T):
            d = ref.drop.astype(np.int32)
            dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
            toward = np.where(np.abs(dx) >= np.abs(dy), np.where(dx > 0, 1, 3), np.where(dy > 0, 2, 0))
            acts = np.where(rng.random((N, A)) < 0.75, toward, rng.integers(0, 9, (N, A))).astype(np.int8)
            draws = rng.random((N, A))
            obs, rew, done, cons, succ = ref.step(acts, draws)
            g_obs, g_rew, g_done, info = env.step(torch.as_tensor(acts.astype(np.int32), device="cuda:0"), draws=draws)
            msg = f"ep{ep} t{t}"
            np.testing.assert_array_equal(_np(env.drop), ref.drop, err_msg=msg + " drop")
            np.testing.assert_array_equal(_np(env.status), ref.status, err_msg=msg + " status")
            np.testing.assert_array_equal(_np(g_obs), obs, err_msg=msg + " obs")
            np.testing.assert_array_equal(_np(env.reward_f64), rew, err_msg=msg + " reward")
            np.testing.assert_array_equal(_np(g_done).astype(np.uint8), done, err_msg=msg + " done")
            np.testing.assert_array_equal(_np(info["constraints"]), cons, err_msg=msg + " punish count")
            np.testing.assert_array_equal(_np(info["success"]), succ, err_msg=msg + " success")
            np.testing.assert_array_equal(_np(env.fails), ref.fails)
            np.testing.assert_array_equal(_np(env.step_count), ref.step_count)
        np.testing.assert_array_equal(_np(env.usage_counts()), ref.usage, err_msg=f"usage ep{ep}")
        np.testing.assert_array_equal(_np(env.get_obs(out=torch.empty_like(env.obs))), ref.observe())


def test_meda_adapter_types_and_generator():
    P = pkg()
    env = P.MEDAEnv(30, 60, 4, fov=19)
    obs = env.reset()
    assert len(obs) == 4 and all(o.dtype == np.float64 and o.shape == (1446,) for o in obs)
    o, r, d, info = env.step([1, 2, 3, 8])
    assert set(r) == {"player_0", "player_1", "player_2", "player_3"} and all(isinstance(v, float) for v in r.values())
    assert info["success"] in (0, 1) and env.get_env_info()["obs_shape"] == (4, 19, 19, 2, 1446)
    env2 = P.MEDAEnv_v0_2(30, 60, 4, fov=19)
    obs = env2.reset()
    assert all(o.dtype == np.int8 and o.shape == (1085,) for o in obs)
    assert env2.get_env_info() == {"n_actions": 9, "n_agents": 4, "obs_shape": (3, 19, 19, 2, 1085), "episode_limit": 90}
    # MEDAEnv_v0_1 (`--version 0.1`, common/config.py:14-16): float64 obs, direction = (dy / width, dx / length)
    g = load_golden("meda_c4")
    env1 = P.MEDAEnv_v0_1(g["W"], g["L"], g["A"], fov=g["fov"], layouts=g["layouts"][0][0])
    obs = env1.reset(layouts=g["layouts"][0][0])
    assert all(o.dtype == np.float64 and o.shape == (1446,) for o in obs)
    want = np.concatenate([g["obs1_reset"][0, 0].astype(np.float64), g["dir1_reset"][0, 0]], axis=-1)
    np.testing.assert_array_equal(np.stack(obs), want)
    assert env1.get_env_info()["obs_shape"] == (4, 19, 19, 2, 1446)
    # info['constraints'] and `fails` are the reference's FLOATS (sums of -0.6 in its accumulation order), bit for bit
    g8 = load_golden("meda_80x80")            # 10 droplets: several close pairs per step
    k = 0
    envg = P.MEDAEnv_v0_2(g8["W"], g8["L"], g8["A"], fov=g8["fov"], layouts=g8["layouts"][0][k])
    envg.reset(layouts=g8["layouts"][0][k])
    fails, n_pun = 0, 0
    for t in range(g8["T"]):
        _, _, _, info = envg.step([int(x) for x in g8["actions"][0, t, k]])
        assert info["constraints"] == g8["constraints"][0, t, k], t
        fails += g8["constraints"][0, t, k]
        assert envg.fails == fails
        n_pun += int(g8["constraints"][0, t, k] != 0)
    assert n_pun > 0
    with pytest.raises(RuntimeError, match="Too many droplets"):
        P.MEDAEnv(30, 60, 9)
    # device task generator: the reference's rejection rules (meda.py:78-81,179-182)
    N, A = 20000, 8
    b = P.BatchedMEDA(N, 30, 60, A, device="cuda:0", seed=5)
    b.reset()
    d = b.drop.to(torch.int32)
    assert int(d[..., 0].min()) >= 2 and int(d[..., 0].max()) <= 57 and int(d[..., 1].min()) >= 2 and int(d[..., 1].max()) <= 27
    for lo in (0, 2):
        p = d[:, :, lo:lo + 2]
        diff = p[:, :, None, :] - p[:, None, :, :]
        d2 = (diff * diff).sum(-1) + torch.eye(A, device="cuda:0", dtype=torch.int32)[None] * 10000
        assert int(d2.min()) >= 81
    own = (d[..., 0] - d[..., 2]).abs().le(4) & (d[..., 1] - d[..., 3]).abs().le(4)
    assert not bool(own.any())
    # auto-reset keeps stepping and resets finished envs
    gen = torch.Generator(device="cuda:0").manual_seed(1)
    for t in range(95):
        acts = torch.randint(0, 9, (N, A), device="cuda:0", generator=gen, dtype=torch.int8)
        obs, rew, done, info = b.step(acts, auto_reset=True)
    assert int(b.step_count.max()) < 90 and int(b.terminated.sum()) == 0


def test_meda_restart_returns_to_the_start_squares():
    P = pkg()
    env = P.BatchedMEDA(500, 30, 60, 4, fov=19, obs_version=2, device="cuda:0", seed=9)
    first = env.reset().clone()
    start = env.drop.clone()
    gen = torch.Generator(device="cuda:0").manual_seed(4)
    for t in range(12):
        env.step(torch.randint(0, 8, (500, 4), device="cuda:0", generator=gen, dtype=torch.int8))
    assert not torch.equal(env.drop, start)
    fails = env.fails.clone()
    obs = env.restart()
    assert torch.equal(env.drop, start) and torch.equal(obs, first)
    assert int(env.step_count.max()) == 0 and int(env.status.max()) == 0 and torch.equal(env.fails, fails)


def test_meda_device_move_draws_equal_injected_philox_draws():
    """Same pin as for DMFB: the MEDA kernel's own draws (Philox4x32-10 keyed by seed, global env, episode, step,
    agent; meda.py:280) equal the host restatement of that stream injected as float64 draws."""
    import philox_ref
    P = pkg()
    N, W, L, A, seed, base = 400, 30, 60, 4, 0xFEDCBA98_76543210, 7_000_000_001
    rng = np.random.default_rng(5)
    kw = dict(fov=19, b_degrade=True, per_degrade=1.0, obs_version=2, device="cuda:0", seed=seed, env_base=base, reward_f64=True)
    a, b = P.BatchedMEDA(N, W, L, A, **kw), P.BatchedMEDA(N, W, L, A, **kw)
    a.reset()
    b.reset()
    assert torch.equal(a.drop, b.drop)
    health = rng.random((N, W, L)) * 0.8 + 0.2
    a.health.copy_(torch.as_tensor(health))
    b.health.copy_(torch.as_tensor(health))
    for t in range(25):
        acts = torch.as_tensor(rng.integers(0, 8, (N, A)).astype(np.int8), device="cuda:0")
        draws = philox_ref.move_draws(seed, base + np.arange(N), _np(b.episode), _np(b.step_count) + 1, A)
        a.step(acts)
        b.step(acts, draws=draws)
        np.testing.assert_array_equal(_np(a.drop), _np(b.drop), err_msg=f"t{t}")
        np.testing.assert_array_equal(_np(a.reward_f64), _np(b.reward_f64), err_msg=f"t{t}")


def test_meda_usage_log_is_transparent():
    """meda_state_t.usage_log: steps log the droplet centres whose footprints addUsage would increment, resets replay
    the log before updateHealth.  Same counters, health and trajectories with and without it - through auto-resets,
    steps past the log capacity and a new chip."""
    P = pkg()
    N, W, L, A = 300, 30, 60, 4
    rng = np.random.default_rng(12)
    kw = dict(fov=19, b_degrade=True, per_degrade=1.0, obs_version=2, device="cuda:0", seed=31, reward_f64=True)
    a = P.BatchedMEDA(N, W, L, A, usage_log=True, **kw)
    b = P.BatchedMEDA(N, W, L, A, usage_log=False, **kw)
    assert a.usage_log is not None and b.usage_log is None and torch.equal(a.drop, b.drop)
    for env in (a, b):
        env.usage.fill_(45)

    def run(steps, tag, **skw):
        for t in range(steps):
            acts = torch.as_tensor(rng.integers(0, 9, (N, A)).astype(np.int8), device="cuda:0")
            a.step(acts, **skw)
            b.step(acts, **skw)
            assert torch.equal(a.drop, b.drop) and torch.equal(a.reward_f64, b.reward_f64), f"{tag} t{t}"
            assert torch.equal(a.health, b.health), f"{tag} t{t} health"
        assert torch.equal(a.usage_counts(), b.usage_counts()), tag
        assert int(a.usage_log_len.max()) == 0

    run(40, "auto-reset", auto_reset=True)
    run(W + L + 12, "past max_step")            # no resets: the log fills up; usage stops growing at max_step anyway
    a.reset()
    b.reset()
    assert torch.equal(a.health, b.health)
    run(30, "after reset", auto_reset=True)


@pytest.mark.parametrize("ver,deg,A,W,L", [(0, False, 4, 30, 60), (2, True, 4, 30, 60), (1, True, 6, 45, 60), (2, False, 10, 80, 80)])
def test_meda_listed_auto_reset_equals_masked_reset(ver, deg, A, W, L):
    """auto_reset is fused into the step kernel (env b); with meda_state_t.reset_list the step only lists the envs that
    terminated and a second kernel resets exactly those (env a); env c steps without auto_reset and then runs a masked
    meda_reset over the envs the step reported as terminated.  Same tasks, observations, health and counters."""
    P = pkg()
    N = 1500
    rng = np.random.default_rng(ver + A)
    kw = dict(fov=19, b_degrade=deg, per_degrade=1.0, obs_version=ver, device="cuda:0", seed=77, reward_f64=True, track_usage=True)
    a = P.BatchedMEDA(N, W, L, A, reset_list=True, **kw)
    b = P.BatchedMEDA(N, W, L, A, reset_list=False, **kw)
    c = P.BatchedMEDA(N, W, L, A, reset_list=False, **kw)
    assert a.reset_list is not None and b.reset_list is None and torch.equal(a.drop, b.drop) and torch.equal(a.drop, c.drop)
    for env in (a, b, c):
        env.usage.fill_(46)
        env.step_count.copy_(torch.arange(N, device="cuda:0", dtype=torch.int32) % env.max_step)   # staggered episodes
    n_resets = 0
    for t in range(W + L + 25):
        d = a.drop.to(torch.int32)
        dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
        toward = torch.where(dx.abs() >= dy.abs(), torch.where(dx > 0, 1, 3), torch.where(dy > 0, 2, 0)).to(torch.int8)
        rnd = torch.as_tensor(rng.integers(0, 9, (N, A)).astype(np.int8), device="cuda:0")
        acts = torch.where(torch.as_tensor(rng.random((N, A)) < 0.8, device="cuda:0"), toward, rnd)
        ep0 = a.episode.clone()
        oa, ra, da, ia = a.step(acts, auto_reset=True)
        ob, rb, db, ib = b.step(acts, auto_reset=True)
        oc, rc_, dc, ic = c.step(acts)
        oc = c.reset(mask=ic["terminated"].to(torch.uint8))
        assert torch.equal(ob, oc) and torch.equal(rb, rc_) and torch.equal(db, dc) and torch.equal(b.drop, c.drop), f"t{t} (fused vs masked)"
        assert torch.equal(b.episode, c.episode) and torch.equal(b.step_count, c.step_count) and torch.equal(b.fails, c.fails)
        n_resets += int((a.episode != ep0).sum())
        assert torch.equal(a.drop, b.drop) and torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), f"t{t}"
        assert torch.equal(a.episode, b.episode) and torch.equal(a.step_count, b.step_count)
        assert torch.equal(a.status, b.status) and torch.equal(a.fails, b.fails) and torch.equal(a.start, b.start)
        assert int(a.terminated.sum()) == 0 and int(a.reset_count.abs().sum()) == 0
        if deg:
            assert torch.equal(a.health, b.health) and torch.equal(b.health, c.health), f"t{t} health"
    assert n_resets > N
    assert torch.equal(a.usage_counts(), b.usage_counts()) and torch.equal(b.usage_counts(), c.usage_counts())


@pytest.mark.parametrize("n_envs,W,L,A,ver,deg", [(500, 30, 60, 4, 0, False), (200, 30, 60, 4, 2, True), (64, 80, 80, 10, 2, False)])
def test_meda_host_buffer_path_equals_device_path(n_envs, W, L, A, ver, deg):
    """meda_host_step / meda_host_reset (host buffers in and out, the reference-facing call shape of MEDAEnv.step,
    meda.py:513-550) return exactly what the device-resident API returns, through fused auto-resets and with the
    CPython-set order table the handle builds for more than 8 droplets."""
    P = pkg()
    rng = np.random.default_rng(n_envs)
    dev = P.BatchedMEDA(n_envs, W, L, A, fov=19, obs_version=ver, b_degrade=deg, per_degrade=1.0, device="cuda:0", seed=31)
    host = P.HostMEDA(n_envs, W, L, A, fov=19, obs_version=ver, b_degrade=deg, per_degrade=1.0, device=0, seed=31)
    o_dev = _np(dev.reset(new_chip=True))
    for _ in range(int(dev.episode[0])):       # tasks are keyed by (seed, env, episode): same number of resets
        o_host = host.reset(new_chip=True)
    np.testing.assert_array_equal(o_host, o_dev)
    for t in range(W + L + 10):
        acts = rng.integers(0, 9, (n_envs, A)).astype(np.int8)
        draws = rng.random((n_envs, A)) if deg else None
        o, r, d, info = dev.step(torch.as_tensor(acts, device="cuda:0"), draws=draws, auto_reset=True)
        ho, hr, hd, hinfo = host.step(acts, draws=draws, auto_reset=True)
        np.testing.assert_array_equal(ho, _np(o), err_msg=f"obs t{t}")
        np.testing.assert_array_equal(hr, _np(r), err_msg=f"reward t{t}")
        np.testing.assert_array_equal(hd, _np(d), err_msg=f"done t{t}")
        np.testing.assert_array_equal(hinfo["constraints"], _np(info["constraints"]))
        np.testing.assert_array_equal(hinfo["success"], _np(info["success"]))
    assert int(dev.episode.max()) > int(dev.episode.min()) or int(dev.episode.max()) > 1   # resets did happen
    host.close()


def test_meda_sub_batch_pipelining_is_transparent():
    """BatchedMEDA(sub_batches=K): K sub-batches on K streams, env for env the single-launch trajectories."""
    P = pkg()
    N, W, L, A = 700, 30, 60, 4
    kw = dict(fov=19, obs_version=2, b_degrade=True, per_degrade=1.0, device="cuda:0", seed=9, reward_f64=True)
    a = P.BatchedMEDA(N, W, L, A, **kw)
    b = P.BatchedMEDA(N, W, L, A, sub_batches=3, **kw)
    assert b._sub is not None and torch.equal(a.drop, b.drop) and torch.equal(a.obs, b.obs)
    gen = torch.Generator(device="cuda:0").manual_seed(4)
    acts = torch.randint(0, 9, (W + L + 12, N, A), device="cuda:0", generator=gen, dtype=torch.int8)
    for t in range(acts.shape[0]):
        a.step(acts[t], auto_reset=True)
        b.step(acts[t], auto_reset=True, join=(t % 7 == 0))
    b.join()
    torch.cuda.synchronize()
    assert torch.equal(a.drop, b.drop) and torch.equal(a.obs, b.obs) and torch.equal(a.reward_f64, b.reward_f64)
    assert torch.equal(a.episode, b.episode) and torch.equal(a.fails, b.fails) and torch.equal(a.status, b.status)
    assert torch.equal(a.health, b.health) and torch.equal(a.usage_counts(), b.usage_counts())
    assert int(a.episode.max()) > 1


@pytest.mark.parametrize("W,L,A", [(30, 60, 4), (80, 80, 10)])
def test_meda_device_tasks_equal_the_restated_generator(W, L, A):
    """meda_generate_tasks against its numpy restatement (tests/layout_ref.py: refresh / addTask / _genLegalDroplet,
    meda.py:161-185,213-233, on one sequential stream per (seed, global env, episode)): after a reset kernel and after
    fused auto-resets inside the step kernel."""
    import layout_ref
    P = pkg()
    N, base, seed = 256, 900, 31
    env = P.BatchedMEDA(N, W, L, A, fov=19, device="cuda:0", seed=seed, env_base=base)
    env.reset()
    want = layout_ref.meda_first_tasks(seed, base + np.arange(N), env.episode.cpu().numpy(), W, L, A)
    assert np.array_equal(env.drop.cpu().numpy(), want)
    assert np.array_equal(env.start.cpu().numpy(), want[:, :, :2])
    epi0 = env.episode.clone()
    gen = torch.Generator(device="cuda:0").manual_seed(4)
    for t in range(W + L + 2):
        env.step(torch.randint(0, 9, (N, A), device="cuda:0", generator=gen, dtype=torch.int8), auto_reset=True)
    assert bool((env.episode > epi0).all())
    want = layout_ref.meda_first_tasks(seed, base + np.arange(N), env.episode.cpu().numpy(), W, L, A)
    assert np.array_equal(env.start.cpu().numpy(), want[:, :, :2])
    assert np.array_equal(env.drop.cpu().numpy()[:, :, 2:], want[:, :, 2:])


def test_meda_device_degrade_matrix_equals_the_restated_draws():
    """MEDAEnv.__init__'s degradation factors (meda.py:494-504) drawn on the device (`reset(new_chip=True)`): bit-equal
    to tests/philox_ref.degrade_matrix (rand * 0.4 + 0.6 in two roundings; 1.0 where rand2 < 1 - per_degrade)."""
    import philox_ref
    P = pkg()
    N, W, L, A, base, seed, per = 48, 30, 60, 4, 300, 99, 0.5
    env = P.BatchedMEDA(N, W, L, A, fov=19, b_degrade=True, per_degrade=per, device="cuda:0", seed=seed, env_base=base)
    env.reset(new_chip=True)
    want = philox_ref.degrade_matrix(seed, base + np.arange(N), env.episode.cpu().numpy(), W * L, per)
    got = env.degrade.cpu().numpy().reshape(N, W * L)
    assert np.array_equal(got, want)
    assert abs(float((got == 1.0).mean()) - (1 - per)) < 0.02 and got.min() >= 0.6
