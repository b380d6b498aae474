"""Host building blocks of the product (compiled into the CPU-only test library) against the reference's C, one stage at
a time: the unstable radix sort's exact permutation, the RMQ chainer, the striped local score."""
import ctypes as C

import numpy as np
import pytest

import chainref
import hostlogic
import kswref


@pytest.fixture(scope="module")
def hl():
    return hostlogic.load()


def test_flag_sort_reproduces_reference_permutation(ref, hl):
    """radix_sort_128x is not stable (ksort.h:101-151): equal keys must come out in the reference's order."""
    rng = np.random.default_rng(2)
    ref.radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
    for n, key_bits in [(0, 8), (1, 8), (2, 1), (64, 3), (65, 3), (300, 2), (5000, 4), (70000, 6), (70000, 40), (200000, 10), (3000, 64)]:
        x = rng.integers(0, 1 << min(key_bits, 62), size=n, dtype=np.uint64)
        if key_bits == 64:
            x = rng.integers(0, 1 << 62, size=n, dtype=np.uint64) << np.uint64(2)
        a = np.stack([x, np.arange(n, dtype=np.uint64)], axis=1).copy()
        b = a.copy()
        ref.radix_sort_128x(a.ctypes.data, a.ctypes.data + 16 * n)
        hl.pgmm_test_flag_sort_128x(C.c_void_p(b.ctypes.data), C.c_size_t(n))
        assert np.array_equal(a, b), (n, key_bits)
        assert np.all(np.diff(a[:, 0].astype(np.float64)) >= 0)
    ref.radix_sort_64.argtypes = [C.c_void_p, C.c_void_p]
    v = rng.integers(0, 1 << 20, size=100000, dtype=np.uint64)
    w = v.copy()
    ref.radix_sort_64(v.ctypes.data, v.ctypes.data + 8 * len(v))
    hl.pgmm_test_flag_sort_64(C.c_void_p(w.ctypes.data), C.c_size_t(len(w)))
    assert np.array_equal(v, w)


def test_chain_rmq_matches_reference(ref, hl):
    """mg_lchain_rmq (lchain.c:250-368) with pangraph's parameters: same chains, same scores, same anchor order."""
    rng = np.random.default_rng(8)
    pen_gap = np.float32(0.8 * 0.01 * 19)
    for n, cap in [(40, 100000), (2000, 100000), (30000, 100000), (30000, 300), (8, 100000)]:
        a = chainref.synth_anchors(rng, n)
        chainref.ref_sort(ref, a)
        want_u, want_a = chainref.ref_chain(ref, a, 10000, 1000, 1000, 25, cap, 3, 40, pen_gap, 0.0)
        mine = a.copy()
        u = np.zeros(len(a) + 1, dtype=np.uint64)
        n_a_out = C.c_int64(0)
        hl.pgmm_test_chain_rmq.restype = C.c_int64
        got_n_u = hl.pgmm_test_chain_rmq(C.c_void_p(mine.ctypes.data), C.c_int64(len(a)), 10000, 1000, 1000, 25, cap, 3, 40,
                                         C.c_float(pen_gap), C.c_float(0.0), C.c_void_p(u.ctypes.data), C.byref(n_a_out))
        assert got_n_u == len(want_u) and n_a_out.value == len(want_a), (n, cap)
        assert np.array_equal(u[:got_n_u], want_u)
        assert np.array_equal(mine[:len(want_a)], want_a)
        if n >= 2000:
            assert len(want_u) > 0
    # a negative score threshold: chain ends with negative scores sort above all others (sign-extended keys, lchain.c:41)
    a = chainref.synth_anchors(rng, 3000, noise=0.5)
    chainref.ref_sort(ref, a)
    want_u, want_a = chainref.ref_chain(ref, a, 10000, 1000, 1000, 25, 100000, 1, -50, pen_gap, 0.0)
    mine = a.copy()
    u = np.zeros(len(a) + 1, dtype=np.uint64)
    n_a_out = C.c_int64(0)
    got_n_u = hl.pgmm_test_chain_rmq(C.c_void_p(mine.ctypes.data), C.c_int64(len(a)), 10000, 1000, 1000, 25, 100000, 1, -50,
                                     C.c_float(pen_gap), C.c_float(0.0), C.c_void_p(u.ctypes.data), C.byref(n_a_out))
    assert got_n_u == len(want_u) and np.array_equal(u[:got_n_u], want_u) and np.array_equal(mine[:len(want_a)], want_a)


def test_ll_local_score_matches_reference(ref, hl):
    """ksw_ll_qinit + ksw_ll_i16 (ksw2_ll_sse.c:37-152): score and both end coordinates."""
    rng = np.random.default_rng(4)
    mat = kswref.simple_mat(1, 9, 1)
    ref.ksw_ll_qinit.restype = C.c_void_p
    ref.ksw_ll_qinit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    ref.ksw_ll_i16.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for i in range(120):
        ql, tl = int(rng.integers(1, 700)), int(rng.integers(1, 700))
        q, t = kswref.random_pair(rng, ql, tl, div=float(rng.choice([0.0, 0.05, 0.4])), indel=0.02, n_frac=float(rng.choice([0, 0.02])),
                                  big_indel=int(rng.choice([0, 25])))
        if i % 7 == 0:  # unrelated
            q = rng.integers(0, 4, size=ql).astype(np.uint8)
        qp = ref.ksw_ll_qinit(None, 2, ql, q.ctypes.data, 5, mat.ctypes.data)
        qe, te = C.c_int(), C.c_int()
        want = ref.ksw_ll_i16(qp, tl, t.ctypes.data, 16, 2, C.byref(qe), C.byref(te))
        libc.free(qp)
        qe2, te2 = C.c_int(), C.c_int()
        got = hl.pgmm_test_ll_score(ql, C.c_void_p(q.ctypes.data), tl, C.c_void_p(t.ctypes.data), C.c_void_p(mat.ctypes.data), 16, 2,
                                    C.byref(qe2), C.byref(te2))
        assert (got, qe2.value, te2.value) == (want, qe.value, te.value), (i, ql, tl)


def test_resident_queries_are_encoded_like_ascii_queries(hl):
    """encode_queries builds the forward and reverse-complement codes of the queries either from ASCII (kNt4) or, for
    pgmm_map_self, from the resident target codes eight bases at a time: same bytes, ambiguous bases and odd lengths included."""
    rng = np.random.default_rng(12)
    seqs = []
    for n in (0, 1, 7, 8, 9, 63, 64, 1000, 4097, 250_001):
        s = np.frombuffer(b"ACGTNacgtnRYK", dtype=np.uint8)[rng.integers(0, 13, size=n)]
        seqs.append(s.tobytes())
    arr = (C.c_char_p * len(seqs))(*seqs)
    lens = (C.c_int * len(seqs))(*[len(s) for s in seqs])
    for threads in (1, 4):
        assert hl.pgmm_test_encode_modes(len(seqs), arr, lens, threads) == 0
