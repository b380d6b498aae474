"""Level-2 sharding and the guide-tree ready-queue with the CUDA mapper: two ranks (two processes, gloo for
the plumbing, both bound to cuda:0 -- the collective itself runs over NCCL in bench.py at N > 1 and over gloo in
tests/test_sharding_gloo.py), each builds the round's index on the GPU, maps its share of the queries through
pgmm_map_batch, rank 0 receives the hit records in query-index order.  Every field must equal what the reference's own C
(oracle/_ref) returns for the serial loop.  Inputs: the first two genomes of the reference's bundled Klebsiella set
(BASELINE config 3) and a four-leaf synthetic tree whose root merge is sharded."""
import os
import pickle

import pytest
import torch
import torch.distributed as dist

from test_sharding_gloo import _init, _spawn

pytestmark = pytest.mark.gpu


def _klebs_worker(rank, world, port, q):
    _init(rank, world, port)
    import realdata
    from pangraph_b200 import abi, sharding
    seqs, _ = realdata.load_pair("klebs")
    names = ["0", "1"]
    idx = abi.Index(seqs, names, "asm10", None, 90)  # replicated index: every rank builds it on its GPU

    def map_fn(ids):
        return [pickle.dumps(h) for h in idx.map_batch([seqs[i] for i in ids], [names[i] for i in ids])]

    got = sharding.map_round_sharded([len(s) for s in seqs], map_fn, torch.device("cpu"))
    mid = idx.mo.mid_occ
    idx.close()
    if rank == 0:
        q.put(([pickle.loads(b) for b in got], mid))
    dist.barrier()
    dist.destroy_process_group()


def test_klebsiella_round_sharded_over_two_ranks(ref):
    import realdata
    from oracle import refmm2
    seqs, _ = realdata.load_pair("klebs")
    want, mid_ref = refmm2.ref_map_all(seqs, ["0", "1"], "asm10", None, 90, threads=8)
    got, mid = _spawn(_klebs_worker, 2, timeout=900)
    assert mid == mid_ref
    assert [len(g) for g in got] == [len(w) for w in want]
    assert got == want


def _leaves():
    from pangraph_b200 import synth
    return [g for _, g in synth.genomes(4, length=200_000, n_rearr=5, len_lo=500, len_hi=15000)]


def _tree_worker(rank, world, port, q):
    _init(rank, world, port)
    from pangraph_b200 import abi, sharding
    gen = _leaves()
    children = [None] * 4 + [(0, 1), (2, 3), (4, 5)]

    def merge(node, left, right, group):
        items = pickle.loads(left)["items"] + pickle.loads(right)["items"]
        seqs, names = [s for _, s in items], [n for n, _ in items]
        idx = abi.Index(seqs, names, "asm10", None, 90)
        if group is None:
            hits = idx.map_batch()
        else:
            got = sharding.map_round_sharded([len(s) for s in seqs], lambda ids: [pickle.dumps(h) for h in idx.map_batch([seqs[i] for i in ids], [names[i] for i in ids])],
                                             torch.device("cpu"), group)
            hits = None if got is None else [pickle.loads(b) for b in got]
        idx.close()
        return pickle.dumps({"items": items, "hits": hits})

    have = sharding.run_tree(children, lambda v: pickle.dumps({"items": [(str(v), gen[v])], "hits": None}), merge, torch.device("cpu"),
                             group=dist.group.WORLD, shard_below=1)
    if rank == 0:
        q.put({v: pickle.loads(b)["hits"] for v, b in have.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_guide_tree_on_two_ranks_root_merge_sharded(ref):
    from oracle import refmm2
    gen = _leaves()
    got = _spawn(_tree_worker, 2, timeout=900)
    for node, ids in ((4, [0, 1]), (5, [2, 3]), (6, [0, 1, 2, 3])):
        want, _ = refmm2.ref_map_all([gen[i] for i in ids], [str(i) for i in ids], "asm10", None, 90, threads=4)
        assert got[node] == want, node
    assert sum(len(h) for h in got[6]) >= 6
