"""Parity at the sizes that are benchmarked: a full 2 x 5 Mbp pair of the bench family, the
first two genomes of the reference's bundled E. coli (BASELINE config 2) and Klebsiella (config 3) sets, and many rounds
in flight at once through the DP service / from many threads sharing one index.  Every field of every mm_reg1_t and
every CIGAR must equal what the reference's own C (oracle/_ref) returns."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compare(seqs, names, preset="asm10", k=None, resident=False):
    from oracle import refmm2
    from pangraph_b200 import abi
    want, mid_ref = refmm2.ref_map_all(seqs, names, preset, k, 90, threads=8)
    abi.get_stats(reset=True)
    if resident:
        idx = abi.Index(seqs, names, preset, k, 90, resident_only=True)
        idx.build()
        got = idx.map_self()
    else:
        idx = abi.Index(seqs, names, preset, k, 90)
        got = idx.map_batch()
    assert idx.mo.mid_occ == mid_ref
    idx.close()
    st = abi.get_stats(reset=True)
    assert [len(g) for g in got] == [len(w) for w in want]
    for qi, (g, w) in enumerate(zip(got, want)):
        for ri, (a, b) in enumerate(zip(g, w)):
            assert a == b, (qi, ri, a[:18], b[:18])
    return sum(len(w) for w in want), st


def test_bench_pair_full_size(ref):
    """Pair 0 of the benchmark's family (2 x 5 Mbp, 1 % divergence, 10 rearrangements each), both call paths."""
    sys.path.insert(0, ROOT)
    import bench
    (seqs, names), = bench.make_pairs(1, 0, 5_000_000)
    n, st = compare(seqs, names)
    assert n >= 20
    n2, _ = compare(seqs, names, resident=True)
    assert n2 == n


def test_ecoli_pair(ref):
    """BASELINE config 2: real repeats, self hits, IS elements; the K4 exits to the host arbiter must be taken ON DEVICE."""
    import realdata
    seqs, _ = realdata.load_pair("ecoli")
    n, st = compare(seqs, ["0", "1"])
    assert n == 2002
    assert st["chain_redo_segments"] > 0 and st["chain_redo_anchors"] > 0
    assert st["anchor_sort_host"] >= 1  # repeats: equal target positions, the host replays the reference's unstable sort


def test_klebsiella_pair(ref):
    import realdata
    seqs, _ = realdata.load_pair("klebs")
    n, st = compare(seqs, ["0", "1"])
    assert n > 1000
    assert st["chain_redo_segments"] > 0


def test_many_rounds_in_flight(ref):
    """12 different rounds from 12 threads at once: their DP waves are merged by the DP service, indices are created and
    destroyed while other rounds' merged waves are running.  Each round's hits must equal the reference's."""
    from oracle import refmm2
    from pangraph_b200 import abi, synth
    rounds = []
    for r in range(12):
        L = 150_000 + 37_000 * (r % 5)
        anc = synth.ancestor(L, 1000 + r)
        if r % 3 == 0:  # repeats: self hits and the host-arbiter exits
            unit = anc[1000:4000].copy()
            for st in (20_000, 70_000, 110_000):
                anc[st:st + len(unit)] = unit
        gs = [synth.mutate(anc, 5000 + 10 * r + i, n_rearr=6, len_lo=500, len_hi=12000).tobytes() for i in range(2 + r % 2)]
        rounds.append((gs, [str(100 * r + i) for i in range(len(gs))]))
    want = [refmm2.ref_map_all(s, n, "asm10", None, 90, threads=4)[0] for s, n in rounds]

    def run(job):
        r, resident = job
        seqs, names = rounds[r]
        if resident:
            idx = abi.Index(seqs, names, "asm10", None, 90, resident_only=True)
            idx.build()
            out = idx.map_self()
        else:
            idx = abi.Index(seqs, names, "asm10", None, 90)
            out = idx.map_batch()
        idx.close()
        return out

    jobs = [(r, rep % 2 == 1) for rep in range(4) for r in range(12)]
    abi.get_stats(reset=True)
    with ThreadPoolExecutor(12) as ex:
        got = list(ex.map(run, jobs))
    for (r, _), g in zip(jobs, got):
        assert g == want[r], r
    st = abi.get_stats(reset=True)
    # small batches take their anchor order from the device when no two anchors of a query share a target position, and from
    # the host's replay of the reference's unstable sort otherwise (the rounds with repeats): both happen here
    assert st["anchor_sort_device"] > 0 and st["anchor_sort_host"] > 0


def test_mm_map_from_many_threads_sharing_one_index(ref):
    """The reference maps from a rayon pool: one index, one mm_tbuf_t per worker, one mm_map per query
    (align_with_minimap2_lib.rs:64-74)."""
    from oracle import refmm2
    from pangraph_b200 import abi, synth
    gs = synth.genomes(16, length=80_000, n_rearr=5, len_lo=300, len_hi=8000)
    seqs, names = [g for _, g in gs], [str(7 * i + 1) for i in range(16)]
    want, _ = refmm2.ref_map_all(seqs, names, "asm10", None, 90, threads=8)
    idx = abi.Index(seqs, names, "asm10", None, 90)
    with ThreadPoolExecutor(16) as ex:
        got = list(ex.map(lambda i: idx.map_one(seqs[i], names[i]), range(16)))
        got2 = list(ex.map(lambda i: idx.map_one(seqs[i], names[i]), range(16)))
    idx.close()
    assert got == want and got2 == want
