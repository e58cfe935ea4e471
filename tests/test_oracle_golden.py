"""Pins the CPU oracle (oracle/*.c) to traces recorded from the unmodified reference env
(tests/golden/*.npz, made by tests/golden/make_golden.py).  Bit-exact everywhere, including the
float64 rewards."""
import numpy as np
import pytest

from conftest import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names("dmfb"))
def test_dmfb_oracle_matches_reference_trace(oracle_lib, name):
    g = load_golden(name)
    K, A, W, L = g["K"], g["A"], g["W"], g["L"]
    nb = int(g.get("n_blocks", 0))
    env = oracle_lib.OracleDMFB(K, W, L, A, fov=g["fov"], stall=bool(g["stall"]), b_degrade=bool(g["b_degrade"]),
                                n_blocks=nb)
    env.degrade[...] = g["degrade"]
    env1 = env.with_version(1)
    obs_t = list(g["obs_t"])
    state_t = list(g["state_t"])

    def check_v01(layers, dirs, msg):
        # DMFBenv_v0_1 (dmfb.py:727-835): int8 layers bit-exact; the reference's float64 direction entries are the
        # emitted numerators divided by length / width (same float64 division)
        o = env1.observe()
        np.testing.assert_array_equal(o[..., :-2], layers, err_msg=msg + " obs v0_1 layers")
        np.testing.assert_array_equal(o[..., -2] / L, dirs[..., 0], err_msg=msg + " obs v0_1 dir y")
        np.testing.assert_array_equal(o[..., -1] / W, dirs[..., 1], err_msg=msg + " obs v0_1 dir x")

    for ep in range(g["n_ep"]):
        obs = env.reset(g["layouts"][ep], new=False, blocks=g["blocks"][ep] if nb else None)
        np.testing.assert_array_equal(obs, g["obs_reset"][ep], err_msg=f"reset obs ep{ep}")
        env1.blocks = env.blocks
        check_v01(g["obs1_reset"][ep], g["dir1_reset"][ep], f"{name} reset ep{ep}")
        np.testing.assert_array_equal(env.health, g["health_reset"][ep], err_msg=f"health at reset ep{ep}")
        np.testing.assert_array_equal(env.usage, g["usage_reset"][ep], err_msg=f"usage at reset ep{ep}")
        for t in range(g["T"]):
            obs, rew, done, cons, succ = env.step(g["actions"][ep, t], g["draws"][ep, t])
            msg = f"{name} ep{ep} t{t}"
            np.testing.assert_array_equal(env.drop[:, :, 0:2], g["pos"][ep, t], err_msg=msg + " pos")
            np.testing.assert_array_equal(rew, g["reward"][ep, t], err_msg=msg + " reward")
            np.testing.assert_array_equal(done, g["done"][ep, t], err_msg=msg + " done")
            np.testing.assert_array_equal(cons, g["constraints"][ep, t], err_msg=msg + " constraints")
            np.testing.assert_array_equal(succ, g["success"][ep, t], err_msg=msg + " success")
            if t in obs_t:
                np.testing.assert_array_equal(obs, g["obs"][ep, obs_t.index(t)], err_msg=msg + " obs")
                check_v01(g["obs1"][ep, obs_t.index(t)], g["dir1"][ep, obs_t.index(t)], msg)
            if t in state_t:
                np.testing.assert_array_equal(env.global_state(), g["state"][ep, state_t.index(t)],
                                              err_msg=msg + " state")
        np.testing.assert_array_equal(env.usage, g["usage_end"][ep], err_msg=f"usage end ep{ep}")
    np.testing.assert_array_equal(env.health, g["health_final"])


@pytest.mark.parametrize("name", golden_names("meda"))
def test_meda_oracle_matches_reference_trace(oracle_lib, name):
    g = load_golden(name)
    K, A, W, L = g["K"], g["A"], g["W"], g["L"]
    env = oracle_lib.OracleMEDA(K, W, L, A, fov=g["fov"], b_degrade=bool(g["b_degrade"]), obs_version=2)
    env0 = env.with_version(0)
    env1 = env.with_version(1)

    def check_v01(o, layers, dirs, msg):
        # MEDAEnv_v0_1 (meda.py:788-844): int8 layers bit-exact; the reference's float64 direction entries are the
        # emitted numerators divided by width / length (same float64 division)
        np.testing.assert_array_equal(o[..., :-2], layers, err_msg=msg + " obs v0_1 layers")
        np.testing.assert_array_equal(o[..., -2] / W, dirs[..., 0], err_msg=msg + " obs v0_1 dir y")
        np.testing.assert_array_equal(o[..., -1] / L, dirs[..., 1], err_msg=msg + " obs v0_1 dir x")

    env.degrade[...] = g["degrade"]
    env.usage[...] = g["usage0"]          # pre-aged chips (meda_*_aged) start with m_usage > 0
    obs_t = list(g["obs_t"])
    for ep in range(g["n_ep"]):
        # the reference updates health AFTER computing the reset observation (meda.py:547-548); the golden
        # health/usage snapshots are taken after reset() returned
        obs2 = env.reset(g["layouts"][ep])
        np.testing.assert_array_equal(obs2, g["obs2_reset"][ep], err_msg=f"v0_2 reset obs ep{ep}")
        np.testing.assert_array_equal(env0.observe(), g["obs0_reset"][ep], err_msg=f"base reset obs ep{ep}")
        check_v01(env1.observe(), g["obs1_reset"][ep], g["dir1_reset"][ep], f"reset ep{ep}")
        np.testing.assert_array_equal(env.health, g["health_reset"][ep], err_msg=f"health at reset ep{ep}")
        np.testing.assert_array_equal(env.usage, g["usage_reset"][ep], err_msg=f"usage at reset ep{ep}")
        for t in range(g["T"]):
            obs, rew, done, cons, succ = env.step(g["actions"][ep, t], g["draws"][ep, t])
            msg = f"{name} ep{ep} t{t}"
            np.testing.assert_array_equal(env.drop[:, :, 0:2], g["pos"][ep, t], err_msg=msg + " pos")
            np.testing.assert_array_equal(rew, g["reward"][ep, t], err_msg=msg + " reward")
            np.testing.assert_array_equal(done, g["done"][ep, t], err_msg=msg + " done")
            np.testing.assert_array_equal(env.status, g["status"][ep, t], err_msg=msg + " status")
            np.testing.assert_allclose(-0.6 * cons, g["constraints"][ep, t], rtol=1e-12, atol=1e-12, err_msg=msg)
            np.testing.assert_array_equal(succ, g["success"][ep, t], err_msg=msg + " success")
            if t in obs_t:
                np.testing.assert_array_equal(obs, g["obs2"][ep, obs_t.index(t)], err_msg=msg + " obs v0_2")
                np.testing.assert_array_equal(env0.observe(), g["obs0"][ep, obs_t.index(t)], err_msg=msg + " obs base")
                check_v01(env1.observe(), g["obs1"][ep, obs_t.index(t)], g["dir1"][ep, obs_t.index(t)], msg)
        np.testing.assert_array_equal(env.usage, g["usage_end"][ep], err_msg=f"usage end ep{ep}")
    np.testing.assert_array_equal(env.health, g["health_final"])
    if name.endswith("_aged"):
        # these traces exist to pin getMoveProb < 1 / failed draws / updateHealth (meda.py:302-309,280,600-605)
        assert g["health_final"].min() < 0.7 and (g["health_final"] < 1.0).sum() > 1000


def test_set_order_emulation_matches_this_cpython(oracle_lib):
    """oracle's emulation of CPython set iteration order (used for MEDAEnv_v0_2 'observed', meda.py:862-872)
    against real python sets built the same way, for every subset of 10 agents."""
    for mask in range(1, 1 << 10):
        s = set()
        for i in range(10):
            if (mask >> i) & 1:
                s.add(i)
        assert oracle_lib.cpython_set_order(mask, 10) == list(s), bin(mask)
