"""GPU parity for the DMFB hot path: CUDA kernels (through the C ABI) vs
 (1) golden traces recorded from the unmodified reference env, and
 (2) the CPU oracle on seeded random inputs at sizes the oracle finishes in seconds,
 (3) size-independent properties at the full 64K-env size.
Bit-exact for every integer/byte tensor and for the float64 rewards; float32 rewards within 1e-6 rel."""
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


def _trace_params():
    """Every reference trace with the usage counters kept (kernel instances <..., DEG=true>) and, for the chips that do
    not degrade, also WITHOUT any degradation state (track_usage=False: health / usage / log all NULL), which is what
    selects the <..., DEG=false> instances that bench.py times (dmfb_kernels.cu StepLaunch)."""
    out = []
    for name in golden_names("dmfb"):
        out.append((name, True))
        if not load_golden(name)["b_degrade"]:
            out.append((name, False))
    return out


@pytest.mark.parametrize("name,track", _trace_params())
def test_dmfb_cuda_matches_reference_trace(name, track):
    g = load_golden(name)
    K, A, W, L = g["K"], g["A"], g["W"], g["L"]
    nb = int(g.get("n_blocks", 0))
    env = pkg().BatchedDMFB(K, W, L, A, nb, fov=g["fov"], stall=bool(g["stall"]), b_degrade=bool(g["b_degrade"]),
                            per_degrade=g["per_degrade"], device="cuda:0", track_usage=track, reward_f64=True,
                            degrade=g["degrade"] if g["b_degrade"] else None, layouts=g["layouts"][0],
                            block_layouts=g["blocks"][0] if nb else None)
    obs_t, state_t = list(g["obs_t"]), list(g["state_t"])
    v01 = pkg().BatchedDMFB(K, W, L, A, nb, fov=g["fov"], device="cuda:0", layouts=g["layouts"][0],
                            block_layouts=g["blocks"][0] if nb else None, obs_version=1)

    def check_v01(layers, dirs, msg):
        # DMFBenv_v0_1 (dmfb.py:723-835) of the same chips: int8 layers bit-exact, direction = numerators / (L, W)
        v01.drop.copy_(env.drop)
        if nb:
            v01.blocks.copy_(env.blocks)
        o = _np(v01.get_obs())
        np.testing.assert_array_equal(o[..., :-2], layers, err_msg=msg + " obs v0_1 layers")
        np.testing.assert_array_equal(o[..., -2] / L, dirs[..., 0], err_msg=msg + " obs v0_1 dir y")
        np.testing.assert_array_equal(o[..., -1] / W, dirs[..., 1], err_msg=msg + " obs v0_1 dir x")

    for ep in range(g["n_ep"]):
        obs = env.reset(new=False, layouts=g["layouts"][ep], block_layouts=g["blocks"][ep] if nb else None)
        np.testing.assert_array_equal(_np(obs), g["obs_reset"][ep], err_msg=f"{name} reset obs ep{ep}")
        check_v01(g["obs1_reset"][ep], g["dir1_reset"][ep], f"{name} reset ep{ep}")
        if g["b_degrade"]:
            np.testing.assert_array_equal(_np(env.health), g["health_reset"][ep], err_msg=f"health ep{ep}")
        if track:
            np.testing.assert_array_equal(_np(env.usage_counts()), g["usage_reset"][ep], err_msg=f"usage ep{ep}")
        else:
            assert env.usage is None and env.health is None and env.usage_log is None
        for t in range(g["T"]):
            acts = torch.as_tensor(g["actions"][ep, t], device="cuda:0")
            obs, rew, done, info = env.step(acts, draws=g["draws"][ep, t])
            msg = f"{name} ep{ep} t{t}"
            np.testing.assert_array_equal(_np(env.positions), g["pos"][ep, t], err_msg=msg + " pos")
            np.testing.assert_array_equal(_np(env.reward_f64), g["reward"][ep, t], err_msg=msg + " reward f64")
            np.testing.assert_allclose(_np(rew), g["reward"][ep, t], rtol=1e-6, atol=0, err_msg=msg + " reward f32")
            np.testing.assert_allclose(_np(info["team_reward"]), g["reward"][ep, t].sum(-1) / A, rtol=1e-6, atol=1e-7)
            np.testing.assert_array_equal(_np(done).astype(np.uint8), g["done"][ep, t], err_msg=msg + " done")
            np.testing.assert_array_equal(_np(info["constraints"]), g["constraints"][ep, t], err_msg=msg + " constraints")
            np.testing.assert_array_equal(_np(info["success"]), g["success"][ep, t], err_msg=msg + " success")
            np.testing.assert_array_equal(_np(info["terminated"]), g["done"][ep, t].all(-1), err_msg=msg + " terminated")
            if t in obs_t:
                np.testing.assert_array_equal(_np(obs), g["obs"][ep, obs_t.index(t)], err_msg=msg + " obs")
                np.testing.assert_array_equal(_np(env.get_obs(out=torch.empty_like(obs))), _np(obs))
                check_v01(g["obs1"][ep, obs_t.index(t)], g["dir1"][ep, obs_t.index(t)], msg)
            if t in state_t:
                np.testing.assert_array_equal(_np(env.get_state()), g["state"][ep, state_t.index(t)],
                                              err_msg=msg + " state")
        if track:
            np.testing.assert_array_equal(_np(env.usage_counts()), g["usage_end"][ep], err_msg=f"usage end ep{ep}")
    if g["b_degrade"]:
        np.testing.assert_array_equal(_np(env.health), g["health_final"])
    assert _np(env.get_avail_actions()).min() == 1
    if name.endswith("_success"):
        assert g["success"].sum() > 100      # these traces exist for the +10/+10 bonus and info['success']


CASES = [
    # N, W, L, A, fov, stall, degrade[, n_blocks[, obs_version]]
    (1000, 10, 10, 4, 9, True, True),
    (400, 14, 14, 4, 9, True, False, 6),
    (300, 20, 24, 7, 7, True, True, 12),
    (200, 12, 12, 2, 5, False, False, 4),
    (333, 20, 20, 10, 9, True, False),
    (130, 50, 50, 10, 9, True, True),
    (257, 12, 15, 6, 7, False, True),
    (64, 16, 11, 3, 8, True, False),
    (97, 30, 30, 7, 11, True, True),
    (40, 25, 40, 5, 13, True, True),
    (33, 9, 9, 2, 9, True, False),
    (31, 40, 40, 12, 19, True, True),
    (1, 10, 10, 4, 9, True, True),
    (20, 128, 128, 32, 19, True, True),          # the library's maxima: chip side, droplets, fov
    (24, 60, 60, 8, 9, True, False, 32),         # ... and obstacles
    (40, 30, 30, 25, 19, True, False),           # odd, long observation rows: 16 of them would not fit shared memory
    (10, 127, 127, 4, 9, True, True),            # odd global-state rows (3 * 127 * 127 bytes)
    (500, 10, 10, 4, 9, True, False, 0, 1),      # DMFBenv_v0_1 through the fused step kernel
    (200, 20, 20, 10, 9, True, True, 0, 1),      # >= 10 droplets: own goal not projected (dmfb.py:756-761)
    (150, 14, 14, 5, 7, True, False, 6, 1),
    (60, 30, 30, 12, 11, False, True, 0, 1),
]


def _case_params():
    # chips that do not degrade also run without any degradation state: the <..., DEG=false> kernel instances
    return [(c, True) for c in CASES] + [(c, False) for c in CASES if not c[6]]


@pytest.mark.parametrize("case,track", _case_params())
def test_dmfb_cuda_matches_oracle_random(oracle_lib, case, track):
    N, W, L, A, fov, stall, deg = case[:7]
    nb = case[7] if len(case) > 7 else 0
    ver = case[8] if len(case) > 8 else 0
    rng = np.random.default_rng(N * 7 + W)
    ref = oracle_lib.OracleDMFB(N, W, L, A, fov=fov, stall=stall, b_degrade=deg, n_blocks=nb, obs_version=ver)
    degrade = rng.random((N, W, L)) * 0.4 + 0.6 if deg else None
    layouts = ref.gen_layouts(seed=N)
    blocks = ref.gen_blocks(N, layouts) if nb else None
    env = pkg().BatchedDMFB(N, W, L, A, nb, fov=fov, stall=stall, b_degrade=deg, per_degrade=1.0, device="cuda:0",
                            track_usage=track, reward_f64=True, degrade=degrade, layouts=layouts, block_layouts=blocks,
                            obs_version=ver)
    if deg:
        ref.degrade[...] = degrade
        # pre-age the chips so that health < 1 matters from the first step
        ref.health[...] = rng.random((N, W, L)) * 0.7 + 0.3
        env.health.copy_(torch.as_tensor(ref.health))
    T = min(2 * (W + L) + 3, 70)
    for ep in range(3):
        layouts = ref.gen_layouts(seed=1000 * ep + N)
        blocks = ref.gen_blocks(1000 * ep + N, layouts) if nb else None
        if ep == 1:  # masked reset: only even envs get a new task
            mask = (np.arange(N) % 2 == 0).astype(np.uint8)
        else:
            mask = None
        o_ref = ref.reset(layouts, mask=mask, blocks=blocks)
        buf = env.obs.clone()
        o_gpu = env.reset(layouts=layouts, mask=mask, block_layouts=blocks)
        sel = slice(None) if mask is None else mask.astype(bool)
        np.testing.assert_array_equal(_np(o_gpu)[sel], o_ref[sel], err_msg=f"reset obs ep{ep}")
        if mask is not None:  # rows of unselected envs must be untouched
            np.testing.assert_array_equal(_np(o_gpu)[~sel], _np(buf)[~sel])
        np.testing.assert_array_equal(_np(env.drop), ref.drop)
        if nb:
            np.testing.assert_array_equal(_np(env.blocks), ref.blocks)
        for t in range(T):
            # goal-biased actions so that arrivals, collisions and the +10/+10 bonus all occur
            d = ref.drop.astype(np.int32)
            dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
            toward = np.where(np.abs(dx) >= np.abs(dy), np.where(dx > 0, 1, 2), np.where(dy > 0, 4, 3))
            toward = np.where((dx == 0) & (dy == 0), 0, toward)
            acts = np.where(rng.random((N, A)) < 0.7, toward, rng.integers(0, 5, (N, A))).astype(np.int8)
            draws = rng.random((N, A))
            obs, rew, done, cons, succ = ref.step(acts, draws)
            g_obs, g_rew, g_done, info = env.step(torch.as_tensor(acts.astype(np.int64), device="cuda:0"), draws=draws)
            msg = f"ep{ep} t{t}"
            np.testing.assert_array_equal(_np(env.drop), ref.drop, err_msg=msg + " drop")
            np.testing.assert_array_equal(_np(g_obs), obs, err_msg=msg + " obs")
            np.testing.assert_array_equal(_np(env.reward_f64), rew, err_msg=msg + " reward")
            np.testing.assert_allclose(_np(g_rew), rew, rtol=1e-6, atol=0)
            np.testing.assert_array_equal(_np(g_done).astype(np.uint8), done, err_msg=msg + " done")
            np.testing.assert_array_equal(_np(info["constraints"]), cons, err_msg=msg + " constraints")
            np.testing.assert_array_equal(_np(info["success"]), succ, err_msg=msg + " success")
            np.testing.assert_array_equal(_np(env.step_count), ref.step_count)
            np.testing.assert_array_equal(_np(env.constraints_cum), ref.constraints)
        if track:
            np.testing.assert_array_equal(_np(env.usage_counts()), ref.usage, err_msg=f"usage ep{ep}")
        np.testing.assert_array_equal(_np(env.get_state()), ref.global_state())
        if deg:
            np.testing.assert_array_equal(_np(env.health), ref.health)


@pytest.mark.parametrize("N,W,L,A,steps", [(65536, 10, 10, 4, 90), (16384, 20, 20, 10, 100)])
def test_benched_instances_with_fused_auto_reset_match_oracle_at_full_size(oracle_lib, N, W, L, A, steps):
    """Exactly what bench.py launches - dmfb_step_kernel<9,4,4,16,false> (C1, 65,536 envs) and <9,10,10,8,false> (C2):
    no degradation state, DMFB_STEP_AUTO_RESET, staggered episode phases - replayed env for env through the oracle.
    The oracle has no task generator of its own in this test: after every step the tasks the DEVICE drew for the envs
    that terminated are read back and injected into the oracle's masked reset (the generator itself is pinned by the
    rejection-rule / uniformity tests).  Every byte of every step's output is compared."""
    fov = 9
    env = pkg().BatchedDMFB(N, W, L, A, fov=fov, stall=True, b_degrade=False, device="cuda:0", seed=1234, reward_f64=True)
    assert env.usage is None and env.health is None and env.usage_log is None      # => DEG = false instance
    ref = oracle_lib.OracleDMFB(N, W, L, A, fov=fov, stall=True, b_degrade=False)
    obs0 = env.reset()
    np.testing.assert_array_equal(ref.reset(_np(env.drop)), _np(obs0))
    phase = (np.arange(N) % env.max_step).astype(np.int32)       # bench.py's staggered episode phases
    env.step_count.copy_(torch.as_tensor(phase))
    ref.step_count[...] = phase
    rng = np.random.default_rng(N + A)
    n_term, n_succ, last_obs = 0, 0, _np(obs0).copy()
    for t in range(steps):
        acts = rng.integers(0, 5, (N, A)).astype(np.int8)
        if t % 3 != 0:   # goal-ward moves two steps out of three, so that episodes also end by success
            d = ref.drop.astype(np.int32)
            dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
            toward = np.where(np.abs(dx) >= np.abs(dy), np.where(dx > 0, 1, 2), np.where(dy > 0, 4, 3))
            acts = np.where((dx == 0) & (dy == 0), 0, toward).astype(np.int8)
        o_ref, r_ref, d_ref, c_ref, s_ref = ref.step(acts)
        o, r, d, info = env.step(torch.as_tensor(acts, device="cuda:0"), auto_reset=True)
        msg = f"t{t}"
        term = d_ref.all(-1)
        np.testing.assert_array_equal(_np(info["terminated"]), term, err_msg=msg + " terminated")
        np.testing.assert_array_equal(_np(env.reward_f64), r_ref, err_msg=msg + " reward")
        np.testing.assert_array_equal(_np(d).astype(np.uint8), d_ref, err_msg=msg + " done")
        np.testing.assert_array_equal(_np(info["constraints"]), c_ref, err_msg=msg + " constraints")
        np.testing.assert_array_equal(_np(info["success"]), s_ref, err_msg=msg + " success")
        np.testing.assert_allclose(_np(info["team_reward"]), r_ref.sum(-1) / A, rtol=1e-6, atol=1e-6)
        drop = _np(env.drop)
        # envs that terminated were reset inside the same launch: hand the oracle the tasks the device drew
        o_reset = ref.reset(drop, mask=term.astype(np.uint8))
        want = np.where(term[:, None, None], o_reset, o_ref)
        np.testing.assert_array_equal(_np(o), want, err_msg=msg + " obs")
        np.testing.assert_array_equal(drop, ref.drop, err_msg=msg + " drop")
        np.testing.assert_array_equal(_np(env.step_count), ref.step_count, err_msg=msg + " step_count")
        np.testing.assert_array_equal(_np(env.constraints_cum), ref.constraints, err_msg=msg + " cum constraints")
        # the new tasks obey _Generate_Start_End's rule (dmfb.py:220) and are new
        if term.any():
            pts = drop[term].astype(np.int32).reshape(-1, A, 2, 2).transpose(0, 2, 1, 3).reshape(-1, 2 * A, 2)
            cheb = np.abs(pts[:, :, None, :] - pts[:, None, :, :]).max(-1) + 9 * np.eye(2 * A, dtype=np.int32)
            assert cheb.min() >= 2, msg
        n_term += int(term.sum())
        n_succ += int(s_ref.sum())
    assert n_term > 2 * N and (n_succ > 0 or A > 4)   # on average every env went through two fused resets


def test_dmfb_known_answers():
    """Constructed layouts with known answers from the reference."""
    P = pkg()

    def run(layout, acts, steps=1, **kw):
        env = P.DMFBenv(10, 10, len(layout), fov=9, layouts=layout, **kw)
        env.reset(layouts=layout)
        out = None
        for _ in range(steps):
            out = env.step(list(acts))
        return out

    # adjacent move: droplet 0 moves right next to droplet 1 which stalls
    obs, r, d, info = run([(2, 2, 8, 2), (4, 2, 4, 9)], [1, 0])
    assert info["constraints"] == 4 and np.allclose([r["player_0"], r["player_1"]], [-4.1, -4.25], rtol=1e-12)
    # move into an occupied cell -> reverted
    obs, r, d, info = run([(2, 2, 8, 2), (3, 2, 3, 9)], [1, 0])
    assert info["constraints"] == 6 and np.allclose([r["player_0"], r["player_1"]], [-6.4, -6.25], rtol=1e-12)
    # wall clamp
    obs, r, d, info = run([(0, 0, 5, 5), (9, 9, 4, 4)], [2, 0])
    assert info["constraints"] == 0 and np.allclose([r["player_0"], r["player_1"]], [-0.4, -0.25], rtol=1e-12)
    # both arrive in the same step without constraints: -0.1 + 10 + 10, success
    obs, r, d, info = run([(1, 1, 2, 1), (7, 7, 7, 8)], [1, 4])
    assert info["success"] == 1 and np.allclose([r["player_0"], r["player_1"]], [19.9, 19.9], rtol=1e-12)
    assert d == {"player_0": True, "player_1": True}
    # the step after everybody is done: 0 + 10 + 10
    obs, r, d, info = run([(1, 1, 2, 1), (7, 7, 7, 8)], [1, 4], steps=2)
    assert np.allclose([r["player_0"], r["player_1"]], [20.0, 20.0], rtol=1e-12)
    assert all(o.dtype == np.int8 and o.shape == (245,) for o in obs)


def test_dmfb_ctor_errors():
    P = pkg()
    with pytest.raises(RuntimeError, match="Fov is too large"):
        P.BatchedDMFB(4, 8, 8, 2, fov=9)
    with pytest.raises(TypeError, match="Too many droplets"):
        P.BatchedDMFB(4, 5, 5, 5, fov=5)
    with pytest.raises(AssertionError):
        P.DMFBenv(4, 10, 2)
    env = P.DMFBenv(10, 10, 4, fov=9)
    with pytest.raises(TypeError):
        env.step([0, 1, 2, 7])
    with pytest.raises(TypeError, match="wrong actions"):
        env.step((0, 1, 2, 3))
    with pytest.raises(RuntimeError):
        env.step([0, 1])


def test_dmfb_freeze_terminated_pads_like_rollout():
    """DMFB_STEP_FREEZE_TERM: a finished env emits the zero padding of rollout.py:131-141."""
    P = pkg()
    layouts = np.array([[(1, 1, 2, 1), (7, 7, 7, 8)], [(1, 1, 9, 9), (7, 7, 0, 0)]], np.uint8)
    env = P.BatchedDMFB(2, 10, 10, 2, fov=9, device="cuda:0", layouts=layouts)
    env.reset(layouts=layouts)
    a = torch.tensor([[1, 4], [0, 0]], device="cuda:0")
    obs, rew, done, info = env.step(a, freeze_terminated=True)
    assert _np(info["terminated"]).tolist() == [True, False]
    pos = _np(env.positions).copy()
    obs, rew, done, info = env.step(a, freeze_terminated=True)
    assert _np(info["padded"]).tolist() == [True, False]
    assert not _np(obs)[0].any() and _np(obs)[1].any()
    assert _np(rew)[0].tolist() == [0.0, 0.0] and _np(env.get_avail_actions())[0].max() == 0
    assert _np(env.get_avail_actions())[1].min() == 1
    np.testing.assert_array_equal(_np(env.positions)[0], pos[0])
    assert int(env.step_count[0]) == 1 and int(env.step_count[1]) == 2


def test_dmfb_device_generator_and_full_size_properties():
    """64K envs (the benchmark size): on-device task generator obeys the reference's rejection rule
    (dmfb.py:220: every pairwise squared distance among the 2A points > 2) and the step keeps the
    invariants the reference guarantees: no two droplets share a cell, positions stay on chip,
    obs layer-0 centre byte equals own index + 1, sharded generation equals the unsharded one."""
    P = pkg()
    N, W, L, A, fov = 65536, 10, 10, 4, 9
    env = P.BatchedDMFB(N, W, L, A, fov=fov, device="cuda:0", seed=1234)
    obs = env.reset()
    d = env.drop.to(torch.int32)
    pts = torch.cat([d[:, :, 0:2], d[:, :, 2:4]], dim=1)  # [N,2A,2]
    diff = pts[:, :, None, :] - pts[:, None, :, :]
    d2 = (diff * diff).sum(-1) + torch.eye(2 * A, device="cuda:0", dtype=torch.int32)[None] * 1000
    assert int(d2.min()) > 2
    assert int(d[:, :, 0].max()) < W and int(d[:, :, 1].max()) < L
    # roughly uniform starts
    hist = torch.bincount((d[:, :, 0] * L + d[:, :, 1]).flatten(), minlength=W * L).float()
    assert hist.min() > 0.5 * hist.mean() and hist.max() < 1.6 * hist.mean()
    # sharded generation == unsharded, env for env
    lo, hi = P.shard_range(N, 1, 4)
    shard = P.BatchedDMFB(hi - lo, W, L, A, fov=fov, device="cuda:0", seed=1234, env_base=lo)
    shard.reset()
    assert torch.equal(shard.drop, env.drop[lo:hi])
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    for t in range(45):
        acts = torch.randint(0, 5, (N, A), device="cuda:0", generator=gen, dtype=torch.int8)
        obs, rew, done, info = env.step(acts)
        p = env.drop[:, :, 0].to(torch.int32) * L + env.drop[:, :, 1].to(torch.int32)
        srt = p.sort(dim=1).values
        assert bool((srt[:, 1:] != srt[:, :-1]).all()), "two droplets share a cell"
        centre = obs.view(N, A, -1)[:, :, (fov // 2) * fov + fov // 2]
        assert torch.equal(centre, torch.arange(1, A + 1, device="cuda:0", dtype=torch.int8).expand(N, A))
    assert bool(done.all()) and int(env.step_count.min()) == 45  # past max_step: dones forced True
    assert int(info["success"].sum()) == 0


@pytest.mark.parametrize("W,L,A,fov,deg", [(10, 10, 4, 9, True), (20, 20, 10, 9, False), (12, 15, 6, 7, True)])
def test_dmfb_fused_auto_reset_equals_step_plus_masked_reset(W, L, A, fov, deg):
    """step(auto_reset=True) == step() followed by reset(mask=terminated) with the same new tasks."""
    P = pkg()
    N = 3000
    kw = dict(fov=fov, b_degrade=deg, per_degrade=1.0, device="cuda:0", seed=77, track_usage=True)
    e1 = P.BatchedDMFB(N, W, L, A, **kw)
    e2 = P.BatchedDMFB(N, W, L, A, **kw)
    assert torch.equal(e1.drop, e2.drop)
    if deg:  # age the chips so that health matters and usage crosses the threshold quickly
        e1.usage.fill_(49); e2.usage.fill_(49)
    gen = torch.Generator(device="cuda:0").manual_seed(3)
    n_resets = 0
    for t in range(2 * (W + L) + 25):
        d = e1.drop.to(torch.int32)
        dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
        toward = torch.where(dx.abs() >= dy.abs(), torch.where(dx > 0, 1, 2), torch.where(dy > 0, 4, 3))
        toward = torch.where((dx == 0) & (dy == 0), 0, toward)
        rnd = torch.randint(0, 5, (N, A), device="cuda:0", generator=gen)
        acts = torch.where(torch.rand(N, A, device="cuda:0", generator=gen) < 0.8, toward, rnd).to(torch.int8)
        draws = torch.rand(N, A, device="cuda:0", generator=gen, dtype=torch.float64)
        o1, r1, d1, i1 = e1.step(acts, draws=draws, auto_reset=True)
        term1 = i1["terminated"].clone()
        o2, r2, d2, i2 = e2.step(acts, draws=draws)
        assert torch.equal(i2["terminated"], term1) and torch.equal(r1, r2) and torch.equal(d1, d2)
        assert torch.equal(i1["success"], i2["success"]) and torch.equal(i1["constraints"], i2["constraints"])
        e2.reset(mask=term1.to(torch.uint8), layouts=e1.drop)
        n_resets += int(term1.sum())
        assert torch.equal(e1.drop, e2.drop) and torch.equal(o1, e2.obs)
        assert torch.equal(e1.step_count, e2.step_count) and torch.equal(e1.constraints_cum, e2.constraints_cum)
        assert torch.equal(e1.terminated, e2.terminated) and int(e1.terminated.sum()) == 0
        assert torch.equal(e1.usage_counts(), e2.usage_counts()) and torch.equal(e1.start, e2.start)
        if deg:
            assert torch.equal(e1.health, e2.health)
    assert n_resets > N  # every env finished at least one episode
    if deg:
        assert float(e1.health.min()) < 1.0


def test_dmfb_device_block_generator_obeys_reference_rules():
    """GenRandomBlocks (dmfb.py:228-251) on the device: blocks inside the chip, never on a start / goal cell, never
    overlapping; auto-reset regenerates them with the task; too many blocks -> none, like the reference."""
    P = pkg()
    N, W, L, A, nb = 20000, 16, 14, 4, 8
    env = P.BatchedDMFB(N, W, L, A, nb, fov=7, device="cuda:0", seed=3)

    def check():
        b = env.blocks.to(torch.int32)                      # [N, nb, 2]
        assert int(b[..., 0].min()) >= 0 and int(b[..., 0].max()) <= W - 4 and int(b[..., 1].max()) <= L - 4
        d = env.drop.to(torch.int32)
        pts = torch.cat([d[:, :, 0:2], d[:, :, 2:4]], dim=1)            # [N, 2A, 2]
        rel = pts[:, :, None, :] - b[:, None, :, :]                      # [N, 2A, nb, 2]
        inside = ((rel >= 0) & (rel <= 1)).all(-1)
        assert not bool(inside.any())
        db = (b[:, :, None, :] - b[:, None, :, :]).abs()
        overlap = (db <= 1).all(-1) & ~torch.eye(nb, dtype=torch.bool, device="cuda:0")[None]
        assert not bool(overlap.any())
    check()
    first = env.blocks.clone()
    env.reset()
    check()
    assert not torch.equal(first, env.blocks)
    gen = torch.Generator(device="cuda:0").manual_seed(2)
    for t in range(2 * (W + L) + 2):
        acts = torch.randint(0, 5, (N, A), device="cuda:0", generator=gen, dtype=torch.int8)
        obs, rew, done, info = env.step(acts, auto_reset=True)
        # no droplet ever stands on a block
        d = env.drop.to(torch.int32)
        rel = d[:, :, None, 0:2] - env.blocks.to(torch.int32)[:, None, :, :]
        assert not bool(((rel >= 0) & (rel <= 1)).all(-1).any())
    check()
    assert P.BatchedDMFB(4, 10, 10, 2, 6, fov=5, device="cuda:0").n_blocks == 0   # 24/100 > 0.2 -> no blocks


def test_empty_and_tiny_batches():
    """N = 0 (empty batch) is a no-op through every entry point; N = 1..3 (ragged: far below one tile) work."""
    P = pkg()
    env = P.BatchedDMFB(0, 10, 10, 4, fov=9, device="cuda:0")
    assert env.reset().shape == (0, 4, 245)
    obs, rew, done, info = env.step(torch.zeros(0, 4, dtype=torch.int8, device="cuda:0"), auto_reset=True)
    assert obs.shape == (0, 4, 245) and rew.shape == (0, 4) and env.get_state().shape == (0, 3, 10, 10)
    assert env.get_obs().shape == (0, 4, 245)
    m = P.BatchedMEDA(0, 30, 60, 4, device="cuda:0")
    assert m.step(torch.zeros(0, 4, dtype=torch.int8, device="cuda:0"))[0].shape == (0, 4, 1085)
    for n in (1, 2, 3):
        e = P.BatchedDMFB(n, 10, 10, 4, fov=9, device="cuda:0", seed=n)
        o = e.reset()
        assert o.shape == (n, 4, 245) and bool((o[:, :, 40] == torch.arange(1, 5, device="cuda:0", dtype=torch.int8)).all())
        for t in range(45):
            o, r, d, info = e.step(torch.randint(0, 5, (n, 4), device="cuda:0"), auto_reset=True)
        assert int(e.step_count.max()) < 40


def test_dmfb_restart_returns_to_the_start_cells():
    P = pkg()
    env = P.BatchedDMFB(700, 10, 10, 4, fov=9, device="cuda:0", seed=2)
    first = env.reset().clone()
    start = env.drop.clone()
    gen = torch.Generator(device="cuda:0").manual_seed(4)
    for t in range(9):
        env.step(torch.randint(0, 5, (700, 4), device="cuda:0", generator=gen, dtype=torch.int8))
    assert not torch.equal(env.drop, start)
    obs = env.restart()
    assert torch.equal(env.drop, start) and torch.equal(obs, first)
    assert int(env.step_count.max()) == 0 and int(env.constraints_cum.max()) == 0


def test_dmfb_v0_1_adapter_returns_reference_types():
    """DMFBenv_v0_1 (`--version 0.1`, common/config.py:6-8): float64 obs of length 4*fov^2+2 with the direction
    ((tar_y - y) / length, (tar_x - x) / width); env_info keeps the BASE obs_shape like the reference."""
    P = pkg()
    g = load_golden("dmfb_c1")
    env = P.DMFBenv_v0_1(g["W"], g["L"], g["A"], fov=g["fov"], layouts=g["layouts"][0][0])
    obs = env.reset(layouts=g["layouts"][0][0])
    assert len(obs) == g["A"] and all(o.dtype == np.float64 and o.shape == (4 * 81 + 2,) for o in obs)
    want = np.concatenate([g["obs1_reset"][0, 0].astype(np.float64), g["dir1_reset"][0, 0]], axis=-1)
    np.testing.assert_array_equal(np.stack(obs), want)
    o, r, d, info = env.step([int(a) for a in g["actions"][0, 0, 0]])
    want = np.concatenate([g["obs1"][0, 0, 0].astype(np.float64), g["dir1"][0, 0, 0]], axis=-1)
    np.testing.assert_array_equal(np.stack(o), want)
    assert env.get_env_info()["obs_shape"] == (3, 9, 9, 2, 245)
    assert env._b.get_env_info()["obs_shape"] == (4, 9, 9, 2, 326)



def test_device_move_draws_equal_injected_philox_draws():
    """Without injected draws the kernel takes random.random() (dmfb.py:335) from Philox4x32-10 keyed by (seed, global
    env, episode, step, agent) and compares it with the health as a 53-bit integer.  The numpy restatement of the same
    stream (tests/philox_ref.py), injected as float64 draws into a twin env, must give the same trajectory - on
    degraded cells, on healthy cells (prob == 1: no draw needed) and on dead cells (prob == 0)."""
    import philox_ref
    P = pkg()
    N, W, L, A, seed, base = 700, 12, 12, 4, 0x12345678_9ABCDEF1, 5_000_000_123   # env index above 2^32
    rng = np.random.default_rng(3)
    kw = dict(fov=9, b_degrade=True, per_degrade=1.0, device="cuda:0", seed=seed, env_base=base, track_usage=True,
              reward_f64=True)
    a, b = P.BatchedDMFB(N, W, L, A, **kw), P.BatchedDMFB(N, W, L, A, **kw)
    a.reset(new=True)
    b.reset(new=True)
    assert torch.equal(a.drop, b.drop)
    health = rng.random((N, W, L))
    health[rng.random((N, W, L)) < 0.2] = 1.0
    health[rng.random((N, W, L)) < 0.05] = 0.0
    a.health.copy_(torch.as_tensor(health))
    b.health.copy_(torch.as_tensor(health))
    moved = 0
    for t in range(30):
        acts = torch.as_tensor(rng.integers(1, 5, (N, A)).astype(np.int8), device="cuda:0")
        draws = philox_ref.move_draws(seed, base + np.arange(N), _np(b.episode), _np(b.step_count) + 1, A)
        before = a.drop.clone()
        a.step(acts)                       # device draws
        b.step(acts, draws=draws)          # the same draws, computed on the host
        np.testing.assert_array_equal(_np(a.drop), _np(b.drop), err_msg=f"t{t}")
        np.testing.assert_array_equal(_np(a.reward_f64), _np(b.reward_f64), err_msg=f"t{t}")
        moved += int((before != a.drop).any(-1).sum())
    assert 0.2 < moved / (30 * N * A) < 0.9   # some moves failed on degraded cells, some succeeded


@pytest.mark.parametrize("n_envs,W,L,A,fov,deg", [(1000, 10, 10, 4, 9, False), (333, 20, 20, 10, 9, True), (64, 12, 15, 6, 7, False)])
def test_host_buffer_path_equals_device_path_plain_and_packed(n_envs, W, L, A, fov, deg):
    """dmfb_host_step (host buffers; what bench.py's e2e times) returns exactly what the device-resident API
    returns, with the plain DMA transfer, with chunk streams, and with the packed transfer (4-bit cells expanded
    by host threads beside the DMA of the rest) at several splits."""
    P = pkg()
    rng = np.random.default_rng(n_envs)
    dev = P.BatchedDMFB(n_envs, W, L, A, fov=fov, b_degrade=deg, per_degrade=1.0, device="cuda:0", seed=77, track_usage=deg)
    hosts = [P.HostDMFB(n_envs, W, L, A, fov=fov, b_degrade=deg, per_degrade=1.0, device=0, seed=77, n_chunks=c)
             for c in (1, 3, 1, 1, 1)]
    hosts[2].set_transfer(4, 50)
    hosts[3].set_transfer(3, 0)        # everything packed
    hosts[4].set_transfer(2, 100)      # pool on, nothing packed
    o_dev = _np(dev.reset(new=True))
    n_resets = int(dev.episode[0])         # tasks are keyed by (seed, env, episode): replay the same episode count
    for h in hosts:
        for _ in range(n_resets):
            o_host = h.reset(new=True)
        np.testing.assert_array_equal(o_host, o_dev)
    for t in range(2 * (W + L) + 5):
        acts = rng.integers(0, 5, (n_envs, A)).astype(np.int8)
        o, r, d, info = dev.step(torch.as_tensor(acts, device="cuda:0"), auto_reset=True)
        for k, h in enumerate(hosts):
            ho, hr, hd, hinfo = h.step(acts, auto_reset=True)
            np.testing.assert_array_equal(ho, _np(o), err_msg=f"host {k} obs t{t}")
            np.testing.assert_array_equal(hr, _np(r), err_msg=f"host {k} reward t{t}")
            np.testing.assert_array_equal(hd, _np(d), err_msg=f"host {k} done t{t}")
            np.testing.assert_array_equal(hinfo["constraints"], _np(info["constraints"]))
            np.testing.assert_array_equal(hinfo["success"], _np(info["success"]))
    for h in hosts:
        h.close()
    with pytest.raises(ValueError):
        big = P.HostDMFB(8, 30, 30, 20, fov=9, device=0)
        try:
            big.set_transfer(2, 50)    # 20 droplets do not fit 4-bit cells
        finally:
            big.close()


def test_device_task_generator_is_uniform_over_the_valid_tasks():
    """_Generate_Start_End (dmfb.py:207-226) redraws the whole point set until no two points are within one cell, i.e.
    it samples UNIFORMLY from the valid ordered point sets.  On a 5x5 chip with one droplet there are 456 valid
    (start, goal) pairs; the device generator must hit only those, all
    equally often (chi-square over ~460K tasks)."""
    P = pkg()
    N, W = 65536, 5
    env = P.BatchedDMFB(N, W, W, 1, fov=5, device="cuda:0", seed=2024)
    counts = torch.zeros(W ** 4, dtype=torch.int64, device="cuda:0")
    reps = 7
    for _ in range(reps):
        env.reset()
        d = env.drop[:, 0].to(torch.int64)
        counts += torch.bincount(((d[:, 0] * W + d[:, 1]) * W + d[:, 2]) * W + d[:, 3], minlength=W ** 4)
    c = counts.cpu().numpy().reshape(W, W, W, W)
    xs = np.arange(W)
    valid = np.maximum(np.abs(xs[:, None, None, None] - xs[None, None, :, None]),
                       np.abs(xs[None, :, None, None] - xs[None, None, None, :])) >= 2
    assert int(valid.sum()) == 456 and int(c[~valid].sum()) == 0
    expected = N * reps / 456
    chi2 = float(((c[valid] - expected) ** 2 / expected).sum())
    assert chi2 < 455 + 5 * np.sqrt(2 * 455), chi2          # 5 sigma above the mean of chi-square(455)
    # two droplets: the four points are pairwise apart, and every cell is used as a start
    env2 = P.BatchedDMFB(N, 6, 7, 2, fov=5, device="cuda:0", seed=7)
    pts = env2.drop.to(torch.int32).reshape(N, 4, 2)
    diff = (pts[:, :, None, :] - pts[:, None, :, :]).abs().amax(-1) + 9 * torch.eye(4, dtype=torch.int32, device="cuda:0")
    assert int(diff.min()) >= 2
    assert len(torch.unique(pts[:, 0, 0] * 7 + pts[:, 0, 1])) == 42


def test_usage_log_is_transparent():
    """Steps append the actuated cells to a per-env log and the resets fold it into m_usage before updateHealth reads
    the counters (dmfb_state_t.usage_log).  With and without the log: same counters, same health, same trajectories -
    through fused auto-resets, masked resets, steps past the log capacity (max_step) and record=False steps."""
    P = pkg()
    N, W, L, A = 700, 10, 12, 4
    rng = np.random.default_rng(8)
    kw = dict(fov=9, b_degrade=True, per_degrade=1.0, device="cuda:0", seed=5, track_usage=True, reward_f64=True)
    a = P.BatchedDMFB(N, W, L, A, usage_log=True, **kw)
    b = P.BatchedDMFB(N, W, L, A, usage_log=False, **kw)
    assert a.usage_log is not None and b.usage_log is None and torch.equal(a.drop, b.drop)
    for env in (a, b):
        env.usage.fill_(47)                     # updateHealth fires soon (usage > 50)

    def run(steps, tag, **skw):
        for t in range(steps):
            acts = torch.as_tensor(rng.integers(0, 5, (N, A)).astype(np.int8), device="cuda:0")
            a.step(acts, **skw)
            b.step(acts, **skw)
            assert torch.equal(a.drop, b.drop) and torch.equal(a.reward_f64, b.reward_f64), f"{tag} t{t}"
            assert torch.equal(a.health, b.health), f"{tag} t{t} health"
        assert torch.equal(a.usage_counts(), b.usage_counts()), tag
        assert int(a.usage_log_len.max()) == 0          # usage_counts() folded everything in

    run(50, "auto-reset", auto_reset=True)
    run(2 * (W + L) + 9, "past the log capacity")       # no reset: the log fills up, the rest is added in place
    mask = (np.arange(N) % 2 == 0).astype(np.uint8)
    a.reset(mask=mask)
    b.reset(mask=mask)
    assert torch.equal(a.health, b.health)
    run(15, "record off", record=False)
    run(30, "after masked reset", auto_reset=True)
    a.reset(new=True)
    b.reset(new=True)
    assert int(a.usage_counts().max()) == 0 and torch.equal(a.health, b.health)


@pytest.mark.parametrize("W,L,A,fov,nb", [(10, 10, 4, 9, 0), (20, 20, 10, 9, 0), (14, 14, 4, 7, 5), (12, 12, 6, 5, 0)])
def test_task_prefetch_is_transparent(W, L, A, fov, nb):
    """dmfb_state_t.next_task / next_cursor: with auto-reset the search for the next episode's task runs ahead, one
    round of attempts per warp and step, and the reset picks the result up (or finishes the search where it stopped).
    The task is the first accepted attempt of (seed, env, episode) either way: trajectories with and without the
    prefetch are identical - through fused resets, explicit generator resets, masked resets with injected tasks (which
    use up an episode number and so invalidate what was prefetched) and restarts."""
    P = pkg()
    N = 3000
    kw = dict(fov=fov, device="cuda:0", seed=99)
    a = P.BatchedDMFB(N, W, L, A, nb, task_prefetch=True, **kw)
    b = P.BatchedDMFB(N, W, L, A, nb, task_prefetch=False, **kw)
    assert a.next_task is not None and b.next_task is None and torch.equal(a.drop, b.drop)
    gen = torch.Generator(device="cuda:0").manual_seed(6)
    ready_seen = 0

    def run(steps, tag):
        nonlocal ready_seen
        for t in range(steps):
            d = a.drop.to(torch.int32)
            dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
            toward = torch.where(dx.abs() >= dy.abs(), torch.where(dx > 0, 1, 2), torch.where(dy > 0, 4, 3))
            rnd = torch.randint(0, 5, (N, A), device="cuda:0", generator=gen)
            acts = torch.where(torch.rand(N, A, device="cuda:0", generator=gen) < 0.6, toward, rnd).to(torch.int8)
            oa, ra, da, ia = a.step(acts, auto_reset=True)
            ob, rb, db, ib = b.step(acts, auto_reset=True)
            assert torch.equal(a.drop, b.drop), f"{tag} t{t}"
            assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(a.step_count, b.step_count), f"{tag} t{t}"
            assert torch.equal(a.episode, b.episode) and torch.equal(a.start, b.start)
            if nb:
                assert torch.equal(a.blocks, b.blocks)
            ready_seen = max(ready_seen, int((a.next_cursor < 0).sum()))      # bit 31 = next task known

    run(2 * (W + L) + 20, "fused resets")
    # the searches do finish ahead of time (the goal-biased policy here ends episodes after ~15 steps and a warp of 8
    # 4-droplet envs advances one search per step, so not every env is ready at any one moment)
    assert ready_seen > (0.75 if (A == 4 and nb == 0) else 0.2) * N
    assert torch.equal(a.reset(), b.reset()) and torch.equal(a.drop, b.drop)          # explicit reset picks them up too
    run(25, "after reset-all")
    mask = (torch.arange(N, device="cuda:0") % 3 == 0).to(torch.uint8)
    lay = a.drop.roll(1, dims=0).clone()
    a.reset(mask=mask, layouts=lay)
    b.reset(mask=mask, layouts=lay)
    run(2 * (W + L) + 5, "after injected tasks")
    a.restart()
    b.restart()
    run(20, "after restart")
    a.check()
    b.check()


@pytest.mark.parametrize("W,L,period,sub", [(20, 20, 1, 1), (20, 20, 3, 1), (20, 20, 4, 4), (18, 32, 2, 1)])
def test_task_search_kernel_is_transparent(W, L, period, sub):
    """Dense 10-droplet chips: the run-ahead search is a kernel of its own behind the step (dmfb_task_search_kernel) -
    32 attempts at once, one per lane, against bit boards in shared memory (chips of at most 30x30 cells; 18x32 takes
    the one-attempt-per-warp flavour), launched behind every `search_period`-th step of a (sub-)batch only
    (DMFB_STEP_SKIP_TASK_SEARCH / DMFB_STEP_SEARCH_SHARE).  Whoever finds it, the task of (seed, env, episode) is the
    first accepted attempt: the trajectories equal those of an env without any prefetch, and the searches do finish
    ahead of the resets."""
    P = pkg()
    N, A = 4096, 10
    kw = dict(fov=9, device="cuda:0", seed=7)
    a = P.BatchedDMFB(N, W, L, A, task_prefetch=True, search_period=period, sub_batches=sub, **kw)
    b = P.BatchedDMFB(N, W, L, A, task_prefetch=False, **kw)
    gen = torch.Generator(device="cuda:0").manual_seed(16)
    picked_up = resets = 0
    for t in range(2 * (W + L) + 30):
        d = a.drop.to(torch.int32)
        dx, dy = d[..., 2] - d[..., 0], d[..., 3] - d[..., 1]
        toward = torch.where(dx.abs() >= dy.abs(), torch.where(dx > 0, 1, 2), torch.where(dy > 0, 4, 3))
        rnd = torch.randint(0, 5, (N, A), device="cuda:0", generator=gen)
        acts = torch.where(torch.rand(N, A, device="cuda:0", generator=gen) < 0.7, toward, rnd).to(torch.int8)
        ready = a.next_cursor < 0                                   # bit 31: the next task is known before the step
        oa, ra, _, ia = a.step(acts, auto_reset=True)
        ob, rb, _, ib = b.step(acts, auto_reset=True)
        picked_up += int((ready & ia["terminated"]).sum())
        resets += int(ia["terminated"].sum())
        assert torch.equal(a.drop, b.drop) and torch.equal(oa, ob) and torch.equal(ra, rb), f"t{t}"
        assert torch.equal(a.episode, b.episode) and torch.equal(ia["terminated"], ib["terminated"])
    assert resets >= N and picked_up > 0.9 * resets, (picked_up, resets)
    a.check()
    b.check()


@pytest.mark.parametrize("W,L,A,fov", [(10, 10, 4, 9), (20, 20, 10, 9), (12, 12, 6, 5), (18, 32, 10, 9)])
def test_device_tasks_are_the_first_accepted_attempt_of_the_stated_stream(W, L, A, fov):
    """The device generator against its numpy restatement (tests/layout_ref.py): the task of (seed, global env, episode)
    is the first attempt of that stream whose 2A cells are pairwise more than one cell apart (dmfb.py:207-226) - after
    an explicit reset (reset kernel), and after fused auto-resets whose tasks were found ahead of time by the run-ahead
    search (inside the step kernel, or by dmfb_task_search_kernel with one attempt per lane / per warp)."""
    import layout_ref
    P = pkg()
    N, base, seed = 768, 4000, 2024
    env = P.BatchedDMFB(N, W, L, A, fov=fov, device="cuda:0", seed=seed, env_base=base)
    env.reset()

    def check(tag):
        want, at = layout_ref.first_accepted_tasks(seed, base + np.arange(N), env.episode.cpu().numpy(), W, L, A)
        assert np.array_equal(env.start.cpu().numpy(), want[:, :, :2]), tag
        assert np.array_equal(env.drop.cpu().numpy()[:, :, 2:], want[:, :, 2:]), tag
        return at

    epi0 = env.episode.clone()
    at = check("reset")
    assert np.array_equal(env.drop.cpu().numpy()[:, :, :2], env.start.cpu().numpy())
    assert at.max() > 3 * at.mean() > 0                      # a geometric number of attempts, not the first one
    gen = torch.Generator(device="cuda:0").manual_seed(3)
    for t in range(2 * (W + L) + 3):
        env.step(torch.randint(0, 5, (N, A), device="cuda:0", generator=gen, dtype=torch.int8), auto_reset=True)
    assert bool((env.episode > epi0).all())
    check("fused auto-reset")
    env.check()


def test_device_degrade_matrices_equal_the_restated_draws():
    """_random_health_statue (dmfb.py:157-164) on the device, `reset(new=True)` without an injected matrix: cell k of
    global env n draws Philox (stream 3) and degrade = rand * 0.4 + 0.6 in two roundings, 1.0 where the second draw is
    below 1 - per_degrade - bit-equal to tests/philox_ref.degrade_matrix, so also independent of sharding."""
    import philox_ref
    P = pkg()
    N, W, L, A, base, seed, per = 96, 12, 15, 4, 7000, (5 << 32) | 77, 0.3
    env = P.BatchedDMFB(N, W, L, A, fov=5, b_degrade=True, per_degrade=per, device="cuda:0", seed=seed, env_base=base)
    env.reset(new=True)
    want = philox_ref.degrade_matrix(seed, base + np.arange(N), env.episode.cpu().numpy(), W * L, per)
    got = env.degrade.cpu().numpy().reshape(N, W * L)
    assert np.array_equal(got, want)
    frac_one = float((got == 1.0).mean())
    assert abs(frac_one - (1 - per)) < 0.02 and got.min() >= 0.6 and got[got < 1.0].max() < 1.0
    assert float(env.health.min()) == 1.0 and int(env.usage_counts().max()) == 0     # refresh(new): a fresh chip


def test_device_blocks_equal_the_restated_generator():
    """GenRandomBlocks (dmfb.py:228-251) on the device against tests/layout_ref.first_blocks: the obstacles of (seed,
    global env, episode) are drawn from their own sequential stream and redrawn while they cover a start / goal cell of
    the env's new task or touch an earlier block - after a reset kernel and after fused auto-resets."""
    import layout_ref
    P = pkg()
    N, W, L, A, nb, base, seed = 256, 14, 14, 4, 5, 123, 808
    env = P.BatchedDMFB(N, W, L, A, nb, fov=7, device="cuda:0", seed=seed, env_base=base)
    env.reset()

    def check(tag):
        epi = env.episode.cpu().numpy()
        tasks, _ = layout_ref.first_accepted_tasks(seed, base + np.arange(N), epi, W, L, A)
        assert np.array_equal(env.start.cpu().numpy(), tasks[:, :, :2]), tag
        want = layout_ref.first_blocks(seed, base + np.arange(N), epi, tasks, W, L, nb)
        assert np.array_equal(env.blocks.cpu().numpy(), want), tag

    check("reset")
    epi0 = env.episode.clone()
    gen = torch.Generator(device="cuda:0").manual_seed(8)
    for t in range(2 * (W + L) + 2):
        env.step(torch.randint(0, 5, (N, A), device="cuda:0", generator=gen, dtype=torch.int8), auto_reset=True)
    assert bool((env.episode > epi0).all())
    check("fused auto-reset")


def test_task_generator_gives_up_recoverably_on_an_impossible_density():
    """8x8 with 9 droplets passes the reference's density check (dmfb.py:144-146) but 18 points that are pairwise not
    within one cell do not fit an 8x8 chip (at most 16 do): the reference would redraw for ever.  The device generator
    gives up, raises DMFB_STATUS_SAMPLER_GAVE_UP (RuntimeError at check time) and leaves the CUDA context usable
    (no __trap)."""
    P = pkg()
    env = P.BatchedDMFB(1, 8, 8, 9, fov=5, device="cuda:0", seed=1)
    with pytest.raises(RuntimeError, match="no legal layout"):
        env.check()
    env.check()                                        # the flag was cleared
    ok = P.BatchedDMFB(64, 10, 10, 4, fov=9, device="cuda:0", seed=1)     # the context is alive
    ok.step(torch.zeros(64, 4, dtype=torch.int8, device="cuda:0"), auto_reset=True)
    ok.check()
    assert int(ok.step_count.max()) == 1


@pytest.mark.parametrize("W,L,A,fov,deg,K", [(10, 10, 4, 9, False, 4), (20, 20, 10, 9, False, 2), (12, 12, 6, 5, True, 3)])
def test_sub_batch_pipelining_is_transparent(W, L, A, fov, deg, K):
    """BatchedDMFB(sub_batches=K) steps the batch as K sub-batches on K streams (pipeline.py).  Env for env the
    trajectories equal the single-launch batch: same tasks (RNG streams follow the global env index), observations,
    rewards, health - with a join after every step, with steps chained without a join, and replayed from a CUDA graph."""
    P = pkg()
    N = 1000                                  # not a multiple of the 64-env sub-batch unit
    kw = dict(fov=fov, b_degrade=deg, per_degrade=1.0, device="cuda:0", seed=5, reward_f64=True)
    a = P.BatchedDMFB(N, W, L, A, **kw)
    b = P.BatchedDMFB(N, W, L, A, sub_batches=K, **kw)
    assert b._sub is not None and len(b._sub.ranges) == K and b._sub.ranges[-1][1] == N
    assert torch.equal(a.drop, b.drop) and torch.equal(a.obs, b.obs)
    gen = torch.Generator(device="cuda:0").manual_seed(3)
    T = 2 * (W + L)
    acts = torch.randint(0, 5, (T + 6, N, A), device="cuda:0", generator=gen, dtype=torch.int8)

    def same(tag):
        b.join()
        torch.cuda.synchronize()
        assert torch.equal(a.drop, b.drop) and torch.equal(a.obs, b.obs), tag
        assert torch.equal(a.reward_f64, b.reward_f64) and torch.equal(a.done, b.done), tag
        assert torch.equal(a.step_count, b.step_count) and torch.equal(a.episode, b.episode), tag
        assert torch.equal(a.constraints, b.constraints) and torch.equal(a.success, b.success), tag
        if deg:
            assert torch.equal(a.health, b.health) and torch.equal(a.usage_counts(), b.usage_counts()), tag

    for t in range(8):                        # joined after every step
        a.step(acts[t], auto_reset=True)
        b.step(acts[t], auto_reset=True)
        same(f"joined t{t}")
    for t in range(8, T + 6):                 # chained: the sub-batch streams run ahead of each other
        a.step(acts[t], auto_reset=True)
        b.step(acts[t], auto_reset=True, join=False)
    same("chained")
    assert int(a.episode.max()) >= 2                        # fused resets did happen
    # the same chain inside a CUDA graph (how bench.py issues it)
    buf_a = torch.zeros(6, N, A, a.D, dtype=torch.int8, device="cuda:0")
    buf_b = torch.zeros_like(buf_a)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for t in range(6):
                b.step(acts[t], auto_reset=True, out=buf_b[t], join=False)
            b.join()
        for t in range(6):
            a.step(acts[t], auto_reset=True, out=buf_a[t])
        g.replay()
        s.synchronize()
    assert torch.equal(buf_a, buf_b) and torch.equal(a.drop, b.drop) and torch.equal(a.reward_f64, b.reward_f64)
    assert torch.equal(b.reset(), a.reset())
