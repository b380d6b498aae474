"""oracle/pgmm_oracle.c::orc_ksw_extd2 (the plain-C restatement) against the reference's ksw_extd2_sse."""
import numpy as np
import pytest

import kswref

PRESETS = {"asm5": (1, 19, 39, 3, 81, 1), "asm10": (1, 9, 16, 2, 41, 1), "asm20": (1, 4, 6, 2, 26, 1)}


def cases(seed, n):
    rng = np.random.default_rng(seed)
    for i in range(n):
        kind = i % 6
        if kind == 0:
            ql, tl, w = int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(0, 50))
        elif kind == 1:
            ql, tl, w = int(rng.integers(150, 320)), int(rng.integers(150, 320)), 150001
        elif kind == 2:
            ql, tl, w = int(rng.integers(200, 900)), int(rng.integers(200, 900)), int(rng.integers(5, 120))
        elif kind == 3:
            ql = tl = int(rng.integers(1, 20)) * 16
            w = int(rng.integers(1, 64))
        elif kind == 4:
            ql, tl, w = int(rng.integers(300, 700)), int(rng.integers(300, 700)), 31
        else:
            ql, tl, w = int(rng.integers(50, 400)), int(rng.integers(50, 400)), 1501
        div = float(rng.choice([0.0, 0.01, 0.05, 0.3]))
        q, t = kswref.random_pair(rng, ql, tl, div=div, indel=float(rng.choice([0.0, 0.01, 0.05])),
                                  n_frac=float(rng.choice([0.0, 0.0, 0.02])),
                                  big_indel=int(rng.choice([0, 0, 30, 150])))
        flag = [kswref.FLAG_LEFT_EXT, kswref.FLAG_FILL1, kswref.FLAG_FILL2, kswref.FLAG_RIGHT_EXT][int(rng.integers(0, 4))]
        preset = list(PRESETS)[int(rng.integers(0, 3))]
        zdrop = int(rng.choice([200, 200, 50, 10]))
        end_bonus = int(rng.choice([-1, -1, 5]))
        yield q, t, w, flag, preset, zdrop, end_bonus


def test_oracle_extd2_matches_reference(ref):
    orc = kswref.load_oracle()
    n = 0
    for q, t, w, flag, preset, zdrop, end_bonus in cases(1234, 600):
        a, b, gq, ge, gq2, ge2 = PRESETS[preset]
        mat = kswref.simple_mat(a, b, 1)
        want = kswref.ref_extd2(ref, q, t, mat, gq, ge, gq2, ge2, w, zdrop, end_bonus, flag)
        got = kswref.orc_extd2(orc, q, t, mat, gq, ge, gq2, ge2, w, zdrop, end_bonus, flag)
        if flag & 0x08:  # approximate-max mode leaves these untouched in the reference (C/ksw2_extd2_sse.c:367-383)
            for k in ("max", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q"):
                assert got[k] == want[k]
        assert got == want, (len(q), len(t), w, hex(flag), preset, zdrop, end_bonus)
        n += 1
    assert n == 600


def test_unbanded_restatement_without_lane_artefacts(ref):
    """When the band cannot bind (every gap fill), evaluating only real cells in plain int arithmetic gives the
    reference's result, and no value ever leaves the signed-byte range (so byte wrap-around is never exercised)."""
    orc = kswref.load_oracle()
    rng = np.random.default_rng(77)
    n = 0
    for i in range(700):
        kind = i % 5
        if kind == 0:
            ql, tl = int(rng.integers(1, 50)), int(rng.integers(1, 50))
        elif kind == 1:
            ql, tl = int(rng.integers(180, 260)), int(rng.integers(180, 260))
        elif kind == 2:
            ql, tl = int(rng.integers(1, 30)), int(rng.integers(200, 900))
        elif kind == 3:
            ql, tl = int(rng.integers(300, 1200)), int(rng.integers(300, 1200))
        else:
            ql = tl = int(rng.integers(1, 30)) * 16
        div = float(rng.choice([0.0, 0.01, 0.1, 0.75]))
        q, t = kswref.random_pair(rng, ql, tl, div=div, indel=float(rng.choice([0.0, 0.02, 0.2])),
                                  n_frac=float(rng.choice([0.0, 0.0, 0.05, 0.5])), big_indel=int(rng.choice([0, 0, 60, 400])))
        if i % 11 == 0:
            q = rng.integers(0, 5, size=ql).astype(np.uint8)  # unrelated, with Ns
        flag = [kswref.FLAG_FILL1, kswref.FLAG_FILL2, kswref.FLAG_RIGHT_EXT, kswref.FLAG_LEFT_EXT][int(rng.integers(0, 4))]
        preset = list(PRESETS)[int(rng.integers(0, 3))]
        a, b, gq, ge, gq2, ge2 = PRESETS[preset]
        mat = kswref.simple_mat(a, b, 1)
        zdrop, end_bonus = int(rng.choice([200, 200, 30])), int(rng.choice([-1, 7]))
        w = max(ql, tl) + int(rng.choice([0, 1, 150001]))
        want = kswref.ref_extd2(ref, q, t, mat, gq, ge, gq2, ge2, w, zdrop, end_bonus, flag)
        got, overflow = kswref.orc_extd2_unbanded(orc, q, t, mat, gq, ge, gq2, ge2, zdrop, end_bonus, flag)
        assert overflow == 0, (i, ql, tl, preset)
        assert got == want, (i, ql, tl, hex(flag), preset)
        n += 1
    assert n == 700
