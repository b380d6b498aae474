"""The N>1 path on CPU: world_size 2 over gloo.  Level 1: rounds / guide-tree merges dealt to ranks, payloads of different
sizes gathered with exact sizes, rank 0 sees every round once and in order.  Level 2: the queries of ONE round sharded over
the ranks against a replicated index, hits back in query-index order and equal to the serial loop (the mapper on CPU is the
reference's own C, oracle/_ref -- the CUDA mapper takes its place in tests/test_gpu_sharded.py)."""
import os
import pickle
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(worker, world, *args, timeout=300):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=timeout)
    for p in procs:
        p.join(timeout=timeout)
        assert p.exitcode == 0
    return res


def _init(rank, world, port):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _gather_worker(rank, world, port, q, n_rounds):
    _init(rank, world, port)
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


@pytest.mark.parametrize("world", [2, 3])
def test_gather_rounds(world):
    n_rounds = 11
    res = _spawn(_gather_worker, world, n_rounds)
    assert [i for i, _, _ in res] == list(range(n_rounds))
    for i, n, first in res:
        assert n == i * 37 % 1000
        if n:
            assert first == bytes([i % 251])


def test_rounds_of_rank_partition():
    from pangraph_b200 import sharding
    for world in (1, 2, 3, 8):
        seen = sorted(r for k in range(world) for r in sharding.rounds_of_rank(17, k, world))
        assert seen == list(range(17))


def test_pack_roundtrip():
    from pangraph_b200 import sharding
    p = [(3, b""), (0, b"abc"), (9, bytes(range(256)))]
    assert sharding.unpack_rounds(sharding.pack_rounds(p)) == p


def test_shard_queries_balanced_and_complete():
    from pangraph_b200 import sharding
    lens = [5_000_000, 4_800_000, 120, 3000, 5_100_000, 4_900_000, 77, 5_050_000, 10]
    for world in (1, 2, 4, 8):
        shards = sharding.shard_queries(lens, world)
        assert sorted(i for s in shards for i in s) == list(range(len(lens)))
        loads = [sum(lens[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lens)  # LPT bound
        assert shards == sharding.shard_queries(lens, world)  # deterministic
    assert sharding.shard_queries([], 3) == [[], [], []]


# ---- level 2: one round, queries sharded ----

def _round_inputs():
    from pangraph_b200 import synth
    gs = synth.genomes(5, length=30_000, n_rearr=4, len_lo=300, len_hi=4000)
    return [g for _, g in gs], [str(11 * i + 3) for i in range(5)]


def _sharded_worker(rank, world, port, q):
    _init(rank, world, port)
    from oracle import refmm2
    from pangraph_b200 import sharding
    seqs, names = _round_inputs()
    lib = refmm2.load_ref()
    idx = refmm2.Index(lib, seqs, names, "asm10", None, 90)  # every rank builds the round's index (replicated)

    def map_fn(ids):
        return [pickle.dumps(idx.map_one(i)) for i in ids]

    got = sharding.map_round_sharded([len(s) for s in seqs], map_fn, torch.device("cpu"))
    idx.close()
    if rank == 0:
        q.put([pickle.loads(b) for b in got])
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_round_sharded_over_two_ranks_equals_serial_loop(ref):
    from oracle import refmm2
    seqs, names = _round_inputs()
    want, _ = refmm2.ref_map_all(seqs, names, "asm10", None, 90)
    got = _spawn(_sharded_worker, 2)
    assert sum(len(w) for w in want) > 0
    assert got == want


# ---- level 1: ready-queue over the guide tree ----

def _kat_tree():
    """(((A,B),(C,D)),(G,H)); of packages/pangraph/src/tree/clade.rs:96-122; node ids: leaves 0..5 = A B C D G H."""
    children = [None] * 6 + [(0, 1), (2, 3), (4, 5), (6, 7), (9, 8)]
    names = ["A", "B", "C", "D", "G", "H", "", "", "", "", ""]
    return children, names


def test_postorder_matches_reference_vector():
    from pangraph_b200 import sharding
    children, names = _kat_tree()
    s = sharding.TreeSchedule(children)
    assert [names[v] for v in s.postorder()] == ["A", "B", "", "C", "D", "", "", "G", "H", "", ""]  # clade.rs:121
    assert s.postorder() == [0, 1, 6, 2, 3, 7, 9, 4, 5, 8, 10]


def test_ready_queue_respects_the_tree():
    from pangraph_b200 import sharding
    children, _ = _kat_tree()
    for world in (1, 2, 4):
        s = sharding.TreeSchedule(children)
        seen, waves = set(range(6)), []
        while not s.finished():
            wave = s.next_wave(world)
            assert wave, "deadlock"
            for v, r in wave:
                assert 0 <= r < world and all(c in seen for c in children[v])
            if world >= len(wave):
                assert len({r for _, r in wave}) == len(wave)  # one merge per rank while ranks are free
            s.complete([v for v, _ in wave])
            seen |= {v for v, _ in wave}
            waves.append(sorted(v for v, _ in wave))
        assert waves == [[6, 7, 8], [9], [10]]
    # a caterpillar tree has no merge-level parallelism: one merge per wave (only level 2 helps there)
    cat = [None] * 4 + [(0, 1), (4, 2), (5, 3)]
    s = sharding.TreeSchedule(cat)
    order = []
    while not s.finished():
        w = s.next_wave(8)
        assert len(w) == 1
        order.append(w[0][0])
        s.complete([w[0][0]])
    assert order == [4, 5, 6]


def _leaf_genomes():
    from pangraph_b200 import synth
    return [g for _, g in synth.genomes(4, length=20_000, n_rearr=3, len_lo=300, len_hi=3000)]


def _merge_stub(refmm2, lib):
    """Alignment half of merge_graphs on a stand-in graph: a node's payload is the list of (name, sequence) below it (no
    reweave here), the merge aligns all of them all-vs-all and records the number of hits."""
    def merge(node, left, right, group):
        from pangraph_b200 import sharding
        items = pickle.loads(left)["items"] + pickle.loads(right)["items"]
        seqs, names = [s for _, s in items], [n for n, _ in items]
        idx = refmm2.Index(lib, seqs, names, "asm10", None, 90)
        if group is None:
            hits = [idx.map_one(i) for i in range(len(seqs))]
        else:  # all ranks on one merge: level 2
            got = sharding.map_round_sharded([len(s) for s in seqs], lambda ids: [pickle.dumps(idx.map_one(i)) for i in ids],
                                             torch.device("cpu"), group)
            hits = None if got is None else [pickle.loads(b) for b in got]
        idx.close()
        return pickle.dumps({"items": items, "hits": hits})
    return merge


def _tree_worker(rank, world, port, q, shard_below):
    _init(rank, world, port)
    from oracle import refmm2
    from pangraph_b200 import sharding
    gen = _leaf_genomes()
    children = [None] * 4 + [(0, 1), (2, 3), (4, 5)]
    lib = refmm2.load_ref()
    have = sharding.run_tree(children, lambda v: pickle.dumps({"items": [(str(v), gen[v])], "hits": None}), _merge_stub(refmm2, lib),
                             torch.device("cpu"), group=dist.group.WORLD if shard_below else None, shard_below=shard_below)
    if rank == 0:
        q.put({v: pickle.loads(b)["hits"] for v, b in have.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shard_below", [0, 1])
def test_tree_over_two_ranks_equals_serial_postorder(ref, shard_below):
    """shard_below=0: merges dealt to ranks (level 1 only).  shard_below=1: waves with fewer ready merges than ranks (the
    root) run with all ranks on one merge and its queries sharded (level 2)."""
    from oracle import refmm2
    gen = _leaf_genomes()
    want = {}
    for node, ids in ((4, [0, 1]), (5, [2, 3]), (6, [0, 1, 2, 3])):
        want[node], _ = refmm2.ref_map_all([gen[i] for i in ids], [str(i) for i in ids], "asm10", None, 90)
    got = _spawn(_tree_worker, 2, shard_below)
    assert {v: got[v] for v in want} == want
    assert sum(len(h) for h in want[6]) > 0
