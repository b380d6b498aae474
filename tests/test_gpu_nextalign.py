"""K6 on the GPU (pgmm_map_variations_batch) against the oracle restatement of the reference's map_variations
(oracle/nextalign_oracle.c, pinned on the reference's unit vectors in tests/test_oracle_nextalign.py): the reference's own
vectors, random related / unrelated pairs with degenerate bands and retries, IUPAC codes, error cases, and block-sized
problems (100 kbp against its consensus) in one batch."""
import numpy as np
import pytest

import naref
from test_nextalign_emul import mutate, rand_seq

pytestmark = pytest.mark.gpu

KEYS = ("subs", "dels", "inss", "hit_boundary", "attempts")


def run(cases, extra=5, attempts=4):
    from pangraph_b200 import abi
    got = abi.map_variations_batch([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], [c[3] for c in cases], extra, attempts)
    for c, g in zip(cases, got):
        want = naref.map_variations(c[0], c[1], c[2], c[3], extra, attempts)
        if isinstance(want, int):
            assert g == want, (c[2], c[3], g, want)
        else:
            assert not isinstance(g, int), (len(c[0]), len(c[1]), c[2], c[3], g)
            assert {k: g[k] for k in KEYS} == {k: want[k] for k in KEYS}, (len(c[0]), len(c[1]), c[2], c[3])
    return got


def test_reference_vectors():
    got = run([("ACTTTGCGTCTGATAGCTTAGCGGATATTTACTGTA", "ACTAGATTGAGTCTGATAGCTTAGCGGATATTGTA", -2, 3),
               ("ACACTGATTTCGTCCCTTAGGTACTCTACACTGTAGCCTA", "CTGATTTAGTCCCTTAGGGGTTACTCTACACTGTAG", 2, 2),
               ("ACACTGATTTCGTCCCTTAGGTACTCTACACTGTAGCCTA", "CCTGACACTGATTTAGTCCTAGGGGTTACTCTACACCGTAGCCTAGCCGCCG", -4, 2),
               ("CGCCCTACTACAAGAGGGAACTTTTTTTTTAAGTATAGCCACAATAGCTGG", "CGCCCTACTACAAGAGGGAACGGGGGGGGGGGGGAAGTATAGCCACAATAGCTGG", -2, 11),
               ("A" * 37, "G" * 18, 70, 0), ("A" * 37, "G" * 18, -70, 0), ("ACGT", "", 0, 0), ("ACGT", "ACxT", 0, 0), ("", "ACGT", 0, 0)])
    assert got[0]["subs"] == [(6, "A")] and got[0]["dels"] == [(29, 4)] and got[0]["inss"] == [(3, "AGA")]  # map_variations.rs:276-280
    assert got[4]["dels"] == [(0, 37)] and got[4]["inss"] == [(37, "G" * 18)]


def test_random_pairs_and_degenerate_bands():
    rng = np.random.default_rng(21)
    for extra, attempts in ((5, 4), (0, 1), (5, 2)):
        cases = []
        for it in range(300):
            ref = rand_seq(rng, int(rng.integers(1, 400)), "ACGT" if it % 7 else "ACGTN")
            kind = it % 5
            if kind == 0:
                qry = rand_seq(rng, int(rng.integers(1, 400)))
            else:
                qry = mutate(rng, ref, iupac=0.02 if kind == 2 else 0.0) or "A"
                if kind == 3:
                    qry = rand_seq(rng, int(rng.integers(0, 30))) + qry[int(rng.integers(0, 20)):]
                if kind == 4:
                    qry = qry[:max(1, len(qry) - int(rng.integers(0, 40)))]
            ms = int(rng.integers(-40, 40)) if it % 3 == 0 else int(rng.integers(-3, 4)) if it % 3 == 1 else int(rng.integers(-500, 500))
            bw = int(rng.integers(0, 6)) if it % 2 else int(rng.integers(0, 120))
            cases.append((ref, qry, ms, bw))
        run(cases, extra, attempts)


def test_block_sized_problems_in_one_batch():
    """What a leaf merge asks for: every node sequence of a block against the block consensus (~100 kbp, 1 % divergence,
    small indels, one larger indel that forces a retry), many problems in one call."""
    from pangraph_b200 import abi
    rng = np.random.default_rng(22)
    cases = []
    for k in range(12):
        ref = rand_seq(rng, 60_000 + 9_000 * k)
        qry = mutate(rng, ref, sub=0.01, indel=0.0005, max_indel=6)
        if k % 3 == 0:
            cut = len(qry) // 2
            qry = qry[:cut] + rand_seq(rng, 60) + qry[cut:]
        cases.append((ref, qry, 0 if k % 2 else -3, 12))
    got = run(cases)
    assert any(g["attempts"] > 1 for g in got)
    for c, g in zip(cases, got):
        assert naref.apply_edit(c[0], g) == c[1]
    _, st = abi.map_variations_batch([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], [c[3] for c in cases], with_stats=True)
    assert st["cells"] > 0 and st["kernel_ms"] > 0


def test_wide_bands():
    rng = np.random.default_rng(23)
    ref = rand_seq(rng, 3000)
    qry = ref[:1200] + rand_seq(rng, 700) + ref[1300:]
    run([(ref, qry, -300, bw) for bw in (40, 200, 900, 2500)], 5, 3)
