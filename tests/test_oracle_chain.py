"""The tree-free restatement of mg_lchain_rmq's score fill (oracle/pgmm_oracle.c::orc_chain_fill) against the product's
host arbiter (the AVL replica that is pinned end to end against the reference's mg_lchain_rmq in test_host_stages.py and
against the committed fixtures in test_golden.py): same f, p, v for every anchor up to the first anchor whose RMQ minimum
is not unique (and the rest of that independent segment) -- the one case where the reference's answer depends on the
shape of its tree."""
import ctypes as C

import numpy as np
import pytest

import chainref
import hostlogic
import kswref


@pytest.fixture(scope="module")
def hl():
    return hostlogic.load()


def orc_fill(orc, a, max_dist, inner, bw, skip, cap, pen_gap, pen_skip):
    n = len(a)
    f, p, v, t = (np.zeros(n + 1, dtype=np.int32) for _ in range(4))
    undet = np.zeros(n + 1, dtype=np.uint8)
    orc.orc_chain_fill.restype = C.c_int64
    orc.orc_chain_fill(C.c_int64(n), C.c_void_p(a.ctypes.data), max_dist, inner, bw, skip, cap, C.c_float(pen_gap),
                       C.c_float(pen_skip), C.c_void_p(f.ctypes.data), C.c_void_p(p.ctypes.data), C.c_void_p(v.ctypes.data),
                       C.c_void_p(t.ctypes.data), C.c_void_p(undet.ctypes.data))
    return f[:n], p[:n], v[:n], undet[:n] == 0


def host_fill(hl, a, max_dist, inner, bw, skip, cap, pen_gap, pen_skip):
    n = len(a)
    f, p, v = (np.zeros(n + 1, dtype=np.int32) for _ in range(3))
    hl.pgmm_test_chain_fill_host(C.c_void_p(a.ctypes.data), C.c_int64(n), max_dist, inner, bw, skip, cap, C.c_float(pen_gap),
                                 C.c_float(pen_skip), C.c_void_p(f.ctypes.data), C.c_void_p(p.ctypes.data), C.c_void_p(v.ctypes.data))
    return f[:n], p[:n], v[:n]


CASES = [dict(), dict(inner=0), dict(bw=100, max_dist=50), dict(skip=0), dict(skip=3, pen_skip=0.05), dict(max_dist=300, inner=100),
         dict(cap=200)]


@pytest.mark.parametrize("params", CASES)
def test_tree_free_fill_equals_the_tree(ref, hl, params):
    orc = kswref.load_oracle()
    rng = np.random.default_rng(77)
    pen_gap = np.float32(0.8 * 0.01 * 19)
    kw = dict(max_dist=5000, inner=1000, bw=1000, skip=25, cap=100000, pen_skip=0.0)
    kw.update(params)
    compared = 0
    for n, noise in [(9, 0.2), (600, 0.2), (6000, 0.4), (12000, 0.1)]:
        a = np.concatenate([chainref.synth_anchors(rng, n, noise=noise), chainref.colinear_anchors(rng, n // 2 + 1, rid=3)])
        chainref.ref_sort(ref, a)
        want = host_fill(hl, a, kw["max_dist"], kw["inner"], kw["bw"], kw["skip"], kw["cap"], pen_gap, kw["pen_skip"])
        f, p, v, ok = orc_fill(orc, a, kw["max_dist"], kw["inner"], kw["bw"], kw["skip"], kw["cap"], pen_gap, kw["pen_skip"])
        assert np.array_equal(f[ok], want[0][ok]) and np.array_equal(p[ok], want[1][ok]) and np.array_equal(v[ok], want[2][ok]), n
        compared += int(ok.sum())
    assert compared > 15000  # segments with a non-unique minimum are the exception
