"""N>1 host logic on CPU: world_size-2 gloo process group (127.0.0.1 rendezvous)."""
import importlib
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("omok-ai_b200.sharding")
    start, count = sh.shard_games(16385, world, rank)
    times, work = sh.reduce_measurements({"gpu_ms": 100.0 + 7 * rank, "e2e_s": 2.0 - rank}, {"sims": 1000 * (rank + 1), "positions": count})
    plies = 3
    boards = torch.full((plies, count % 5 + 2, 81), rank, dtype=torch.uint8)
    policy = torch.full((plies, count % 5 + 2, 81), float(rank))
    status = torch.full((plies, count % 5 + 2), rank, dtype=torch.int8)
    merged = sh.gather_replay(boards, policy, status, dst=0)
    q.put((rank, start, count, times, work, None if merged is None else [tuple(m.shape) for m in merged],
           None if merged is None else merged[0][0, :, 0].tolist(), sh.rank_seed(5, rank)))
    dist.destroy_process_group()


def test_two_rank_partition_reduce_and_replay_merge():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, c0, t0, w0, shp0, col0, seed0), (r1, s1, c1, t1, w1, shp1, col1, seed1) = res
    assert (s0, c0, s1, c1) == (0, 8193, 8193, 8192)  # contiguous, exhaustive, balanced
    assert t0 == t1 == {"gpu_ms": 107.0, "e2e_s": 2.0}  # max over ranks
    assert w0 == w1 == {"sims": 3000, "positions": 16385}  # sum over ranks
    n0, n1 = c0 % 5 + 2, c1 % 5 + 2
    assert shp1 is None and shp0 == [(3, n0 + n1, 81), (3, n0 + n1, 81), (3, n0 + n1)]
    assert col0 == [0] * n0 + [1] * n1  # rank blocks concatenated along the game axis, in rank order
    assert seed0 != seed1


def test_single_process_passthrough():
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("omok-ai_b200.sharding")
    assert sh.shard_games(10, 1, 0) == (0, 10)
    assert [sh.shard_games(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    t, w = sh.reduce_measurements({"a": 1.5}, {"b": 2})
    assert t == {"a": 1.5} and w == {"b": 2}
