"""The numpy restatement of the device task generator (tests/layout_ref.py) obeys the reference's rule itself."""
import numpy as np
import pytest

import layout_ref


@pytest.mark.parametrize("W,L,A", [(10, 10, 4), (20, 20, 10), (9, 9, 2)])
def test_restated_generator_obeys_generate_start_end(W, L, A):
    """_Generate_Start_End (dmfb.py:207-226): 2A cells on the chip, min pairwise squared distance > 2; whole-set
    rejection means a geometric number of attempts, and different (env, episode) keys give different tasks."""
    n = 200
    tasks, at = layout_ref.first_accepted_tasks(77, 10 + np.arange(n), 1 + np.arange(n) % 3, W, L, A)
    pts = tasks.reshape(n, 2 * A, 2).astype(np.int64)          # (x, y) of start_0, goal_0, start_1, ...
    assert pts[..., 0].max() < W and pts[..., 1].max() < L and pts.min() >= 0
    d2 = ((pts[:, :, None, :] - pts[:, None, :, :]) ** 2).sum(-1)
    d2[:, np.arange(2 * A), np.arange(2 * A)] = 99
    assert d2.min() > 2                                        # dmfb.py:220
    assert at.min() >= 0 and at.max() > at.mean()
    again, at2 = layout_ref.first_accepted_tasks(77, 10 + np.arange(n), 1 + np.arange(n) % 3, W, L, A)
    assert np.array_equal(tasks, again) and np.array_equal(at, at2)
    other, _ = layout_ref.first_accepted_tasks(78, 10 + np.arange(n), 1 + np.arange(n) % 3, W, L, A)
    assert (tasks != other).any(axis=(1, 2)).mean() > 0.95


def test_mix64_known_answers():
    """splitmix64's finaliser (Steele, Lea, Flood 2014): outputs of the reference implementation for state 0 -
    mix64(k * 0x9E3779B97F4A7C15) for k = 1, 2, 3."""
    with np.errstate(over="ignore"):
        z = layout_ref.mix64(np.arange(1, 4, dtype=np.uint64) * layout_ref.PHI)
    assert [int(v) for v in z] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]


def test_restated_meda_generator_obeys_gen_legal_droplet():
    """meda.py:213-233: centres in [r, dim-r-1] (x against `length`, y against `width`), starts pairwise at least 9
    apart, destinations too, and no destination overlapping its own start (|dx| <= 4 and |dy| <= 4)."""
    W, L, A, r = 30, 60, 4, 2
    t = layout_ref.meda_first_tasks(9, 100 + np.arange(64), 1 + np.arange(64) % 2, W, L, A).astype(np.int64)
    for a, b in ((0, 1), (2, 3)):
        assert t[..., a].min() >= r and t[..., a].max() <= L - r - 1          # x
        assert t[..., b].min() >= r and t[..., b].max() <= W - r - 1          # y
        d2 = (t[:, :, None, a] - t[:, None, :, a]) ** 2 + (t[:, :, None, b] - t[:, None, :, b]) ** 2
        d2[:, np.arange(A), np.arange(A)] = 999
        assert d2.min() >= 81
    assert not ((np.abs(t[..., 2] - t[..., 0]) <= 2 * r) & (np.abs(t[..., 3] - t[..., 1]) <= 2 * r)).any()


def test_restated_block_generator_obeys_gen_random_blocks():
    """dmfb.py:228-251: 2x2 blocks inside the chip that cover no start / goal cell and do not overlap each other."""
    W, L, A, nb, n = 14, 14, 4, 5, 64
    tasks, _ = layout_ref.first_accepted_tasks(3, np.arange(n), np.ones(n), W, L, A)
    blocks = layout_ref.first_blocks(3, np.arange(n), np.ones(n), tasks, W, L, nb).astype(np.int64)
    assert blocks.min() >= 0 and blocks[..., 0].max() <= W - 4 and blocks[..., 1].max() <= L - 4
    pts = tasks.reshape(n, 2 * A, 2).astype(np.int64)
    dx = pts[:, None, :, 0] - blocks[:, :, None, 0]
    dy = pts[:, None, :, 1] - blocks[:, :, None, 1]
    assert not ((dx >= 0) & (dx <= 1) & (dy >= 0) & (dy <= 1)).any()
    bx = np.abs(blocks[:, :, None, 0] - blocks[:, None, :, 0])
    by = np.abs(blocks[:, :, None, 1] - blocks[:, None, :, 1])
    touch = (bx <= 1) & (by <= 1)
    touch[:, np.arange(nb), np.arange(nb)] = False
    assert not touch.any()
