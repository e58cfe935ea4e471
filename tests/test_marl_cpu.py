"""CPU checks of the batched callers of the hot path: CRNN parity with the reference network,
replay-buffer ring semantics, VDN learner, and the 2-rank (gloo) gradient all-reduce."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

OBS_SHAPE = (3, 9, 9, 2, 245)


@pytest.fixture(scope="module")
def P():
    return importlib.import_module("marl-dmfb_b200")


def test_crnn_parameter_count_and_names(P):
    net = P.CRNN(OBS_SHAPE, 5, 128, 24)
    assert sum(p.numel() for p in net.parameters()) == 290765      # SURVEY section 2 row 9 (probe of the reference CRNN)
    assert set(net.state_dict()) == {f"{m}.{w}" for m in ("conv1", "conv2", "mlp1", "fc1") for w in ("weight", "bias")} | \
        {"rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh"}
    net19 = P.CRNN((3, 19, 19, 2, 1085), 9, 128, 24)              # fov 19: stride-2 conv, then the SAME conv3 twice
    q, h = net19(torch.zeros(2, 1085 + 9), torch.zeros(2, 128))
    assert q.shape == (2, 9) and h.shape == (2, 128)


@pytest.mark.skipif(not os.path.isdir("/root/reference/network"), reason="reference checkout not mounted")
def test_crnn_matches_reference_network(P):
    sys.path.insert(0, "/root/reference")
    try:
        from network.base_net import CRNN as RefCRNN
    finally:
        sys.path.pop(0)

    class Args:
        obs_shape, hyper_hidden_dim, rnn_hidden_dim, n_actions, fov = OBS_SHAPE, 24, 128, 5, 9
    torch.manual_seed(0)
    ref = RefCRNN(Args)
    mine = P.CRNN(OBS_SHAPE, 5, 128, 24)
    mine.load_state_dict(ref.state_dict())                            # same parameter names / shapes
    x = torch.randint(0, 5, (7, 250)).float()
    h = torch.randn(7, 128)
    q_ref, h_ref = ref(x, h)
    q, h2 = mine(x, h)
    assert torch.equal(q, q_ref) and torch.equal(h2, h_ref)


@pytest.mark.skipif(not os.path.isdir("/root/reference/network"), reason="reference checkout not mounted")
def test_qmix_networks_match_reference(P):
    sys.path.insert(0, "/root/reference")
    try:
        from network.base_net import RNN as RefRNN
        from network.qmix_net import QMixNet as RefMix
    finally:
        sys.path.pop(0)

    class Args:
        state_shape, n_agents, qmix_hidden_dim, hyper_hidden_dim, rnn_hidden_dim, n_actions = 300, 4, 32, 64, 64, 5
        two_hyper_layers = False
    for two in (False, True):
        Args.two_hyper_layers = two
        torch.manual_seed(0)
        ref = RefMix(Args)
        mine = P.QMixNet(300, 4, 32, 64, two)
        mine.load_state_dict(ref.state_dict())                        # same parameter names / shapes
        q, s = torch.randn(3, 7, 4), torch.randn(3, 7, 300)
        assert torch.equal(ref(q, s), mine(q, s))
    ref = RefRNN(254, Args)
    mine = P.RNN(254, 5, 64)
    mine.load_state_dict(ref.state_dict())
    x, h = torch.randn(6, 254), torch.randn(6, 64)
    assert all(torch.equal(a, b) for a, b in zip(ref(x, h), mine(x, h)))


def test_qmix_mixer_is_monotonic_and_learner_reduces_td_error(P):
    mix = P.QMixNet(300, 4)
    s = torch.randn(5, 3, 300)
    q = torch.randn(5, 3, 4)
    bump = torch.zeros_like(q)
    bump[..., 2] = 0.5
    assert bool((mix(q + bump, s) >= mix(q, s)).all())                # dQtot/dQa >= 0 (abs of the hyper weights)
    learner = P.QMIXLearner(OBS_SHAPE, 4, 5, 300, "cpu", seed=1)
    batch = _synthetic_batch(6, 5, 4, 245, 5, seed=2)
    g = torch.Generator().manual_seed(5)
    s_all = torch.randint(0, 5, (6, 6, 300), generator=g, dtype=torch.int8)
    batch["s"], batch["s_next"] = s_all[:, :5], s_all[:, 1:]
    losses = [float(learner.learn(dict(batch), step)) for step in range(12)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    ep = P.EpisodeBatch(2, 4, 4, 245, 5, "cpu", state_dim=300)
    d = ep.as_dict()
    assert d["s"].shape == (2, 4, 300) and d["s_next"].shape == (2, 4, 300)
    buf = P.ReplayBufferGPU(3, 4, 4, 245, 5, "cpu", state_dim=300)
    ep.s_all.fill_(7)
    buf.store_episodes(ep)
    assert int(buf.sample(4)["s"].max()) == 7


def _synthetic_batch(B, T, A, D, n_act, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    o = torch.randint(0, 5, (B, T + 1, A, D), generator=g, dtype=torch.int8)
    u = torch.randint(0, n_act, (B, T, A, 1), generator=g, dtype=torch.int8)
    onehot = torch.nn.functional.one_hot(u[..., 0].long(), n_act).to(torch.int8)
    r = torch.randn(B, T, 1, generator=g)
    padded = torch.zeros(B, T, 1, dtype=torch.bool)
    term = torch.zeros(B, T, 1, dtype=torch.bool)
    term[:, -1] = True
    avail = torch.ones(B, T + 1, A, n_act, dtype=torch.int8)
    return {"o": o[:, :T], "o_next": o[:, 1:], "u": u, "r": r, "avail_u": avail[:, :T], "avail_u_next": avail[:, 1:],
            "u_onehot": onehot, "padded": padded, "terminated": term}


def test_vdn_learner_reduces_td_error(P):
    learner = P.VDNLearner(OBS_SHAPE, 4, 5, "cpu", seed=1)
    batch = _synthetic_batch(6, 5, 4, 245, 5, seed=2)
    assert learner.max_episode_len(batch) == 5
    losses = [float(learner.learn(dict(batch), step)) for step in range(12)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


def test_max_episode_len_follows_reference_rule(P):
    term = torch.zeros(3, 6, 1, dtype=torch.bool)
    term[0, 2:] = True      # first terminated index 2 -> len 3
    term[1, 4:] = True      # -> len 5
    assert P.VDNLearner.max_episode_len({"terminated": term}) == 5   # episode 2 never terminates -> contributes 0


def test_replay_buffer_ring_and_sampling(P):
    T, A, D, n_act = 4, 2, 11, 5
    buf = P.ReplayBufferGPU(5, T, A, D, n_act, "cpu", seed=0)
    for k in range(4):                                  # 4 batches of 2 episodes into a ring of 5
        ep = P.EpisodeBatch(2, T, A, D, n_act, "cpu")
        ep.r.fill_(float(k))
        buf.store_episodes(ep)
    # reference _get_storage_idx: 0-1, 2-3, then [4, 0] (wrap), then 1-2
    assert buf.current_size == 5 and buf.current_idx == 3
    assert buf.r[0, :, 0].tolist() == [2.0, 3.0, 3.0, 1.0, 2.0]
    s = buf.sample(64)
    assert s["o"].shape == (64, T, A, D) and s["o_next"].shape == (64, T, A, D) and s["u"].dtype == torch.int8
    assert set(s["r"][:, 0, 0].tolist()) <= {1.0, 2.0, 3.0}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ragged_batch():
    """4 episodes of 3 transitions; the two episodes of rank 1 are shorter (1 and 2 live transitions), so the two ranks
    hold different numbers of valid transitions (6 vs 3)."""
    full = _synthetic_batch(4, 3, 4, 245, 5, seed=9)
    for b, n in ((2, 1), (3, 2)):
        full["padded"][b, n:] = True
        full["terminated"][b, n - 1:] = True
        full["r"][b, n:] = 0
    return full


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = importlib.import_module("marl-dmfb_b200")
    learner = P.VDNLearner(OBS_SHAPE, 4, 5, "cpu", seed=3, world_size=world)
    full = _ragged_batch()
    mine = {k: v[rank * 2:(rank + 1) * 2] for k, v in full.items()}     # 2 episodes per rank
    learner.learn(mine, 0)
    flat = torch.cat([p.detach().flatten() for p in learner.eval_rnn.parameters()])
    q.put((rank, flat.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_equals_large_batch(P):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks end with identical parameters ...
    np.testing.assert_array_equal(out[0], out[1])
    # ... equal to one process learning on the concatenated batch, although the ranks hold different numbers of valid
    # transitions: every rank normalises by the GLOBAL count (marl.global_mask_sum)
    single = P.VDNLearner(OBS_SHAPE, 4, 5, "cpu", seed=3)
    single.learn(_ragged_batch(), 0)
    flat = torch.cat([p.detach().flatten() for p in single.eval_rnn.parameters()]).numpy()
    np.testing.assert_allclose(out[0], flat, rtol=2e-4, atol=2e-6)
