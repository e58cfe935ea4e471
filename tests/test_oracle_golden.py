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
    env = oracle_lib.OracleDMFB(K, W, L, A, fov=g["fov"], stall=bool(g["stall"]), b_degrade=bool(g["b_degrade"]))
    env.degrade[...] = g["degrade"]
    obs_t = list(g["obs_t"])
    state_t = list(g["state_t"])
    for ep in range(g["n_ep"]):
        obs = env.reset(g["layouts"][ep], new=False)
        np.testing.assert_array_equal(obs, g["obs_reset"][ep], err_msg=f"reset obs ep{ep}")
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
            if t in state_t:
                np.testing.assert_array_equal(env.global_state(), g["state"][ep, state_t.index(t)],
                                              err_msg=msg + " state")
        np.testing.assert_array_equal(env.usage, g["usage_end"][ep], err_msg=f"usage end ep{ep}")
    np.testing.assert_array_equal(env.health, g["health_final"])
