"""World-size-2 (gloo, CPU) check of the multi-GPU host logic: ranks agree on a disjoint, exhaustive env
partition, `env_base` offsets are consistent, and the max-over-ranks timing reduction bench.py relies on works."""
import importlib
import os
import socket

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = importlib.import_module("marl-dmfb_b200")
    lo, hi = P.shard_range(n_total, rank, world)
    spans = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(spans, torch.tensor([lo, hi]))
    owned = torch.zeros(n_total, dtype=torch.int32)
    owned[lo:hi] += 1
    dist.all_reduce(owned)                       # every env owned exactly once
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # bench.py: time = max over ranks
    q.put((rank, [s.tolist() for s in spans], int(owned.min()), int(owned.max()), float(t)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [131072, 1001])
def test_two_rank_env_partition(n_total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, spans, mn, mx, tmax in out:
        assert spans[0][0] == 0 and spans[-1][1] == n_total and spans[0][1] == spans[1][0]
        assert (mn, mx) == (1, 1)
        assert tmax == 11.0
