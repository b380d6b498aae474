"""oracle/pgmm_oracle.c::orc_sketch against the reference's mm_sketch (and its hash against the reference's own known
answers, packages/pangraph/src/distance/mash/hash.rs:20-27)."""
import ctypes as C

import numpy as np

import kswref


def orc_sketch(lib, seq, w, k, rid=0):
    seq = seq if isinstance(seq, bytes) else seq.encode()
    x = np.zeros(len(seq) + 4, dtype=np.uint64)
    y = np.zeros(len(seq) + 4, dtype=np.uint64)
    lib.orc_sketch.restype = C.c_long
    n = lib.orc_sketch(seq, len(seq), w, k, rid, C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data))
    return list(zip((int(v) for v in x[:n]), (int(v) for v in y[:n])))


def test_oracle_sketch_matches_reference(ref):
    from oracle import refmm2
    orc = kswref.load_oracle()
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for w, k in [(19, 19), (10, 19), (19, 10), (10, 10), (5, 15), (1, 7), (50, 28), (255, 4), (3, 2)]:
        for n in (1, k - 1, k, w + k - 2, w + k - 1, w + k, 500, 6000):
            if n <= 0:
                continue
            s = acgt[rng.integers(0, 4, size=n)].copy()
            if n > 100:
                s[rng.random(n) < 0.01] = ord("N")
                s[50:60] = ord("n")
            if n == 500:
                s[100:400] = np.tile(s[100:137], 9)[:300]  # a tandem repeat: many identical k-mers in a window
            b = s.tobytes()
            assert orc_sketch(orc, b, w, k, 3) == refmm2.ref_sketch(ref, b, w, k, 3), (w, k, n)
        assert orc_sketch(orc, b"ACGT" * 50, w, k) == refmm2.ref_sketch(ref, b"ACGT" * 50, w, k)
        assert orc_sketch(orc, b"acgtnACGTU" * 30, w, k) == refmm2.ref_sketch(ref, b"acgtnACGTU" * 30, w, k)
