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
