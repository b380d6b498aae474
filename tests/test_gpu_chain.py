"""K4 (CUDA chain score fill) + the host backtrack through the C-ABI against the reference's mg_lchain_rmq: same
chains, same scores, same anchors in the same order."""
import numpy as np
import pytest

import chainref

pytestmark = pytest.mark.gpu

PEN_GAP = np.float32(0.8 * 0.01 * 19)


def check(ref, a, max_dist=5000, inner=1000, bw=1000, skip=25, cap=100000, min_cnt=3, min_sc=40, pen_skip=0.0, host_redo=True):
    from pangraph_b200 import abi
    chainref.ref_sort(ref, a)
    want_u, want_a = chainref.ref_chain(ref, a, max_dist, inner, bw, skip, cap, min_cnt, min_sc, PEN_GAP, pen_skip)
    u, kept, fpv, seg = abi.chain_rmq(a, max_dist, inner, bw, skip, cap, min_cnt, min_sc, PEN_GAP, pen_skip, host_redo=host_redo)
    assert np.array_equal(u, want_u)
    assert np.array_equal(kept, want_a)
    return len(want_u), seg


@pytest.mark.parametrize("n,cap", [(0, 100000), (1, 100000), (8, 100000), (40, 100000), (2000, 100000), (30000, 100000), (30000, 300)])
def test_noisy_anchor_sets(ref, n, cap):
    """Collinear runs with jumps, 20 % noise anchors and repeated target positions; both strands, two targets."""
    rng = np.random.default_rng(8 + n)
    if n >= 8:
        a = chainref.synth_anchors(rng, n)
    else:
        a = np.array([[1000, (19 << 32) | 700]] * n, dtype=np.uint64).reshape(n, 2)
    n_u, seg = check(ref, a, max_dist=10000, cap=cap)
    if n >= 2000:
        assert n_u > 0


@pytest.mark.parametrize("params", [dict(), dict(inner=0), dict(bw=100, max_dist=50), dict(skip=0), dict(skip=3, pen_skip=0.05),
                                    dict(max_dist=300, inner=100)])
def test_parameter_corners(ref, params):
    rng = np.random.default_rng(31)
    a = np.concatenate([chainref.synth_anchors(rng, 6000, noise=0.4), chainref.colinear_anchors(rng, 5000, rid=3)])
    check(ref, a, **params)


def test_genome_like_anchors_stay_on_the_device(ref):
    """A long collinear stretch with substitutions, small indels and rearrangement jumps -- the benchmark's shape.
    No segment may need the host arbiter (no equal priorities on such data)."""
    rng = np.random.default_rng(5)
    a = np.concatenate([chainref.colinear_anchors(rng, 200000), chainref.colinear_anchors(rng, 30000, strand=1, y0=40000)])
    n_u, seg = check(ref, a, host_redo=False)
    assert n_u > 5 and seg[1] == 0


def test_dense_repeat_windows_go_to_the_host_arbiter(ref):
    """An RMQ window beyond the device limit (every query position hits every target position of a repeat array)."""
    from pangraph_b200 import abi
    xs = np.repeat(np.arange(1000, 1000 + 3 * 150, 3, dtype=np.uint64), 60)
    ys = np.tile(np.arange(500, 500 + 7 * 60, 7, dtype=np.uint64), 150)
    a = np.stack([xs, (np.uint64(19) << np.uint64(32)) | ys], axis=1).copy()
    n_u, seg = check(ref, a)
    assert seg[1] >= 1


def test_chain_fixture_without_the_reference_library():
    """K4 + backtrack on the committed mg_lchain_rmq fixtures (tests/golden/golden_chain.npz): no oracle/_ref needed."""
    import os

    from pangraph_b200 import abi
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_chain.npz"))
    p = g["params"]
    for name in ("noisy", "colinear", "repeat_array"):
        u, kept, fpv, seg = abi.chain_rmq(g[name + "_in"], int(p[0]), int(p[1]), int(p[2]), int(p[3]), int(p[4]), int(p[5]), int(p[6]),
                                          np.float32(p[7]), float(p[8]))
        assert np.array_equal(u, g[name + "_u"]) and np.array_equal(kept, g[name + "_kept"]), name


def test_fill_arrays_against_the_tree_free_oracle():
    """f, p, v of every anchor as K4 fills them (segments it hands back are filled by the host arbiter) against
    oracle/pgmm_oracle.c::orc_chain_fill, wherever the oracle's answer is determined (unique RMQ minima)."""
    import ctypes as C

    import kswref
    from pangraph_b200 import abi
    orc = kswref.load_oracle()
    orc.orc_chain_fill.restype = C.c_int64
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "golden_chain.npz"))
    rng = np.random.default_rng(99)
    sets = [g["noisy_in"], g["colinear_in"], g["repeat_array_in"]]
    big = chainref.colinear_anchors(rng, 60000)
    sets.append(big[np.argsort(big[:, 0], kind="stable")])
    checked = 0
    for a in sets:
        a = np.ascontiguousarray(a)
        n = len(a)
        u, kept, fpv, seg = abi.chain_rmq(a, 10000, 1000, 1000, 25, 100000, 3, 40, PEN_GAP, 0.0)
        f, p, v, t = (np.zeros(n + 1, dtype=np.int32) for _ in range(4))
        undet = np.zeros(n + 1, dtype=np.uint8)
        orc.orc_chain_fill(C.c_int64(n), C.c_void_p(a.ctypes.data), 10000, 1000, 1000, 25, 100000, C.c_float(PEN_GAP), C.c_float(0.0),
                           C.c_void_p(f.ctypes.data), C.c_void_p(p.ctypes.data), C.c_void_p(v.ctypes.data), C.c_void_p(t.ctypes.data),
                           C.c_void_p(undet.ctypes.data))
        ok = undet[:n] == 0
        assert np.array_equal(fpv[0][ok], f[:n][ok]) and np.array_equal(fpv[1][ok], p[:n][ok]) and np.array_equal(fpv[2][ok], v[:n][ok])
        checked += int(ok.sum())
    assert checked > 60000
