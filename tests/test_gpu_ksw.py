"""K5 (CUDA ksw_extd2) through the C-ABI against the reference's ksw_extd2_sse and the plain-C oracle."""
import numpy as np
import pytest

import kswref
from test_oracle_ksw import PRESETS, cases

pytestmark = pytest.mark.gpu

EZ = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end")


def run_batch(problems, preset, rev=False, budget=0):
    from pangraph_b200 import abi
    a, b, gq, ge, gq2, ge2 = PRESETS[preset]
    qs, ts, qo, to = [], [], [], []
    qpos = tpos = 0
    for q, t, *_ in problems:
        qs.append(q[::-1] if rev else q), ts.append(t[::-1] if rev else t)
        qo.append(qpos), to.append(tpos)
        qpos += len(q) + 3
        tpos += len(t) + 5
    qbuf = np.full(qpos + 8, 7, dtype=np.uint8)
    tbuf = np.full(tpos + 8, 7, dtype=np.uint8)
    for (q, t, *_), o1, o2, qq, tt in zip(problems, qo, to, qs, ts):
        qbuf[o1:o1 + len(q)] = qq
        tbuf[o2:o2 + len(t)] = tt
    flags = [p[3] | (0x10000 if rev else 0) for p in problems]
    ez, cigs, ms = abi.ksw_extd2_batch([len(p[0]) for p in problems], [len(p[1]) for p in problems], qo, to, qbuf, tbuf,
                                       [p[2] for p in problems], [p[4] for p in problems], [p[5] for p in problems], flags,
                                       a, b, 1, gq, ge, gq2, ge2, arena_budget_bytes=budget)
    return ez, cigs, ms


def check(problems, preset, ref, rev=False, budget=0):
    a, b, gq, ge, gq2, ge2 = PRESETS[preset]
    mat = kswref.simple_mat(a, b, 1)
    ez, cigs, _ = run_batch(problems, preset, rev, budget)
    bad = []
    for i, (q, t, w, flag, zdrop, end_bonus) in enumerate(problems):
        want = kswref.ref_extd2(ref, q, t, mat, gq, ge, gq2, ge2, w, zdrop, end_bonus, flag)
        got = {k: int(ez[i, j]) for j, k in enumerate(EZ)}
        got["cigar"] = cigs[i]
        if got != want:
            bad.append((i, len(q), len(t), w, hex(flag), zdrop, end_bonus,
                        {k: (got[k], want[k]) for k in want if got[k] != want[k] and k != "cigar"},
                        got["cigar"][:6], want["cigar"][:6]))
    assert not bad, f"{len(bad)} of {len(problems)} differ; first: {bad[:3]}"


@pytest.mark.parametrize("preset", ["asm5", "asm10", "asm20"])
def test_ksw_random_problems_match_reference(ref, preset):
    probs = [(q, t, w, flag, zd, eb) for q, t, w, flag, p, zd, eb in cases(99, 900) if p == preset]
    assert len(probs) > 200
    check(probs, preset, ref)


def test_ksw_reversed_windows(ref):
    probs = [(q, t, w, flag, zd, eb) for q, t, w, flag, p, zd, eb in cases(5, 240) if p == "asm10"]
    check(probs, "asm10", ref, rev=True)


def test_ksw_long_banded_extensions(ref):
    """The 4k-16k class: band 1501 (align.c:590), z-drop, extension-only, both gap alignments; several CTA sizes."""
    rng = np.random.default_rng(7)
    probs = []
    for i in range(24):
        ql, tl = int(rng.integers(1500, 7000)), int(rng.integers(1500, 7000))
        q, t = kswref.random_pair(rng, ql, tl, div=float(rng.choice([0.01, 0.03, 0.2])), indel=0.005,
                                  big_indel=int(rng.choice([0, 0, 400, 2500])))
        flag = [kswref.FLAG_LEFT_EXT, kswref.FLAG_RIGHT_EXT, kswref.FLAG_FILL2, kswref.FLAG_FILL1][i % 4]
        w = 1501 if flag != kswref.FLAG_FILL1 else 150001
        probs.append((q, t, w, flag, 200, -1))
    check(probs, "asm10", ref)


def test_ksw_global_scratch_and_small_arena(ref):
    """A target too long for shared memory takes the global-memory state path; a tiny arena forces several waves."""
    rng = np.random.default_rng(11)
    probs = []
    q, t = kswref.random_pair(rng, 900, 21000, div=0.02, indel=0.002)
    probs.append((q, t, 1501, kswref.FLAG_RIGHT_EXT, 200, -1))
    q, t = kswref.random_pair(rng, 18000, 17500, div=0.02, indel=0.002)
    probs.append((q, t, 1501, kswref.FLAG_FILL2, 200, -1))
    for _ in range(40):
        q, t = kswref.random_pair(rng, int(rng.integers(100, 400)), int(rng.integers(100, 400)), div=0.05)
        probs.append((q, t, 150001, kswref.FLAG_FILL1, 200, -1))
    check(probs, "asm10", ref, budget=8 << 20)


def test_ksw_edge_cases(ref):
    """Empty windows, 1x1, all-N, identical and unrelated sequences."""
    rng = np.random.default_rng(3)
    one = np.array([2], dtype=np.uint8)
    probs = [(one, one, 1501, kswref.FLAG_RIGHT_EXT, 200, -1), (one, np.array([1], dtype=np.uint8), 150001, kswref.FLAG_FILL1, 200, -1)]
    same = rng.integers(0, 4, size=333).astype(np.uint8)
    probs.append((same, same.copy(), 150001, kswref.FLAG_FILL1, 200, -1))
    probs.append((same, rng.integers(0, 4, size=200).astype(np.uint8), 1501, kswref.FLAG_FILL2, 200, -1))
    probs.append((np.full(64, 4, dtype=np.uint8), np.full(80, 4, dtype=np.uint8), 1501, kswref.FLAG_LEFT_EXT, 200, -1))
    probs.append((same[:16], same[:160], 3, kswref.FLAG_FILL2, 200, -1))
    probs.append((same[:160], same[:16], 3, kswref.FLAG_RIGHT_EXT, 50, 5))
    check(probs, "asm10", ref)
    from pangraph_b200 import abi
    ez, cigs, _ = abi.ksw_extd2_batch([0, 5], [7, 0], [0, 0], [0, 0], same, same, [10, 10], [200, 200], [-1, -1], [0, 0],
                                      1, 9, 1, 16, 2, 41, 1)
    for i in range(2):  # ksw_reset_extz only (ksw2_extd2_sse.c:70-71)
        assert list(ez[i]) == [0, 0, -1, -1, kswref_neg(), -1, kswref_neg(), -1, kswref_neg(), 0, 0]
        assert cigs[i] == ()


def kswref_neg():
    return -0x40000000


def test_small_fill_kernel_k5a(ref):
    """K5a (16-bit lanes, real cells only) takes the first-pass gap fills with tlen <= 256, qlen <= 1024 and a band that
    cannot bind: every small shape, the size limits, long queries, Ns, all presets."""
    rng = np.random.default_rng(21)
    probs = {p: [] for p in PRESETS}
    for ql in range(1, 14):
        for tl in range(1, 14):
            q = rng.integers(0, 4, size=ql).astype(np.uint8)
            t = rng.integers(0, 4, size=tl).astype(np.uint8)
            probs[list(PRESETS)[(ql + tl) % 3]].append((q, t, 150001, kswref.FLAG_FILL1, 200, -1))
    shapes = [(256, 256), (255, 256), (1024, 256), (1023, 3), (1, 256), (256, 1), (700, 64), (31, 33), (32, 32), (64, 63), (65, 64),
              (200, 201), (205, 231), (300, 129), (128, 127)]
    for i, (ql, tl) in enumerate(shapes * 3):
        q, t = kswref.random_pair(rng, ql, tl, div=float(rng.choice([0.0, 0.02, 0.3])), indel=float(rng.choice([0.0, 0.03])),
                                  n_frac=float(rng.choice([0.0, 0.03])), big_indel=int(rng.choice([0, 0, 40])))
        w = max(ql, tl) + int(rng.choice([0, 5, 150001]))
        probs[list(PRESETS)[i % 3]].append((q, t, w, kswref.FLAG_FILL1, 200, -1))
    for preset, ps in probs.items():
        check(ps, preset, ref)


def test_wide_fill_kernel_k5b(ref):
    """K5b takes every problem whose band cannot bind that K5a does not: first and second (exact, z-drop) passes of the
    long fills across inversions / big indels.  Shapes around the four width configurations, early z-drops, Ns."""
    rng = np.random.default_rng(31)
    probs = []
    shapes = [(3, 2), (40, 255), (256, 256), (257, 257), (300, 1024), (1025, 1025), (1100, 700), (2000, 2100), (4096, 4096), (4097, 3000),
              (5000, 6000), (900, 8192), (8000, 8192), (1200, 31), (64, 2049)]
    for i, (ql, tl) in enumerate(shapes * 2):
        div = float(rng.choice([0.01, 0.05, 0.3]))
        q, t = kswref.random_pair(rng, ql, tl, div=div, indel=0.01, n_frac=float(rng.choice([0.0, 0.01])),
                                  big_indel=int(rng.choice([0, 0, 300])))
        if i % 5 == 0 and ql > 600:  # an unrelated stretch in the middle: the exact pass z-drops inside it
            a = ql // 3
            q[a:a + ql // 3] = rng.integers(0, 4, size=ql // 3)
        flag = kswref.FLAG_FILL2 if i % 2 else kswref.FLAG_FILL1
        probs.append((q, t, max(ql, tl) + int(rng.choice([0, 150001])), flag, int(rng.choice([200, 200, 60])), -1))
    check(probs, "asm10", ref)
    check([p for p in probs if len(p[0]) <= 2100], "asm5", ref)
    check([p for p in probs if len(p[0]) <= 2100], "asm20", ref)
