"""The N>1 path on CPU: world_size 2 over gloo.  Rounds are dealt by index, payloads of different sizes are gathered,
rank 0 sees every round exactly once and in round order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rounds, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pangraph_b200 import sharding
    mine = sharding.rounds_of_rank(n_rounds, rank, world)
    payloads = [(r, bytes([r % 251]) * (r * 37 % 1000)) for r in mine]  # round r -> r*37%1000 bytes of value r
    got = sharding.gather_rounds(payloads, torch.device("cpu"))
    if rank == 0:
        q.put([(i, len(b), b[:1]) for i, b in got])
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rounds_world2():
    world, n_rounds = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port_shared, n_rounds, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert [i for i, _, _ in res] == list(range(n_rounds))
    for i, n, first in res:
        assert n == i * 37 % 1000
        if n:
            assert first == bytes([i % 251])


_free_port_shared = _free_port()


def test_rounds_of_rank_partition():
    from pangraph_b200 import sharding
    for world in (1, 2, 3, 8):
        seen = sorted(r for k in range(world) for r in sharding.rounds_of_rank(17, k, world))
        assert seen == list(range(17))


def test_pack_roundtrip():
    from pangraph_b200 import sharding
    p = [(3, b""), (0, b"abc"), (9, bytes(range(256)))]
    assert sharding.unpack_rounds(sharding.pack_rounds(p)) == p
