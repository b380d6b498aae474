"""Shared by the chaining tests: synthetic anchor sets and the reference's own mg_lchain_rmq / radix_sort_128x
(oracle/_ref/libmm2ref.so) as the checker."""
import ctypes as C

import numpy as np


def synth_anchors(rng, n, span=19, noise=0.2, n_ref=2):
    """Anchors of a few collinear runs plus noise, with repeated target positions (ties), as the anchor sort leaves them."""
    xs, ys = [], []
    for rid in range(n_ref):
        for strand in (0, 1):
            m = n // (2 * n_ref)
            pos = np.cumsum(rng.integers(1, 40, size=m))
            q = pos + rng.integers(-3, 4, size=m) + int(rng.integers(0, 5000))
            jump = int(rng.integers(0, max(m, 1)))
            q[jump:] += int(rng.integers(-3000, 3000))
            noise_idx = rng.random(m) < noise
            q[noise_idx] = rng.integers(0, int(pos[-1]) + 6000, size=int(noise_idx.sum()))
            dup = rng.random(m) < 0.05
            pos[1:][dup[1:]] = pos[:-1][dup[1:]]
            xs.append((np.uint64(strand) << np.uint64(63)) | (np.uint64(rid) << np.uint64(32)) | pos.astype(np.uint64))
            ys.append((np.uint64(span) << np.uint64(32)) | np.clip(q, span, None).astype(np.uint64))
    a = np.stack([np.concatenate(xs), np.concatenate(ys)], axis=1).copy()
    return a



def colinear_anchors(rng, n, span=19, div=0.01, x0=1000, y0=500, rid=0, strand=0):
    """What a pair of related genomes gives: minimizer matches every ~10 bp on one diagonal, interrupted at substitutions,
    with small indels shifting the diagonal and a few large jumps (rearrangement borders)."""
    step = rng.integers(1, 20, size=n)
    gap = rng.random(n) < div * 8  # a substitution wipes out the k-mers over it
    step[gap] += span + rng.integers(0, 30, size=int(gap.sum()))
    x = x0 + np.cumsum(step)
    shift = np.zeros(n, dtype=np.int64)
    ind = rng.random(n) < 0.004
    shift[ind] = rng.integers(-3, 4, size=int(ind.sum()))
    big = rng.random(n) < 0.0005
    shift[big] = rng.integers(-20000, 20000, size=int(big.sum()))
    y = y0 + np.cumsum(step) + np.cumsum(shift)
    y = np.clip(y, span, None)
    hi = (np.uint64(strand) << np.uint64(63)) | (np.uint64(rid) << np.uint64(32))
    return np.stack([hi | x.astype(np.uint64), (np.uint64(span) << np.uint64(32)) | y.astype(np.uint64)], axis=1).copy()


def ref_sort(ref, a):
    ref.radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
    ref.radix_sort_128x.restype = None
    ref.radix_sort_128x(a.ctypes.data, a.ctypes.data + 16 * len(a))


def ref_chain(ref, a, max_dist, max_dist_inner, bw, max_skip, cap, min_cnt, min_sc, pen_gap, pen_skip):
    """(u, kept anchors) from the reference's mg_lchain_rmq on sorted anchors a[n,2] (uint64)."""
    ref.mg_lchain_rmq.restype = C.c_void_p
    ref.mg_lchain_rmq.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int64,
                                  C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_void_p]
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    if len(a) == 0:
        return np.zeros(0, np.uint64), np.zeros((0, 2), np.uint64)
    buf = libc.malloc(16 * len(a))  # the reference frees its input: hand it a malloc()ed copy
    C.memmove(buf, a.ctypes.data, 16 * len(a))
    n_u, u_ptr = C.c_int(0), C.c_void_p()
    out = ref.mg_lchain_rmq(max_dist, max_dist_inner, bw, max_skip, cap, min_cnt, min_sc, C.c_float(pen_gap), C.c_float(pen_skip),
                            len(a), buf, C.byref(n_u), C.byref(u_ptr), None)
    u = np.ctypeslib.as_array(C.cast(u_ptr, C.POINTER(C.c_uint64)), (n_u.value,)).copy() if n_u.value else np.zeros(0, np.uint64)
    n_a = int((u & np.uint64(0xffffffff)).sum())
    kept = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint64)), (n_a, 2)).copy() if n_a else np.zeros((0, 2), np.uint64)
    if out:
        libc.free(out)
    if u_ptr.value:
        libc.free(u_ptr.value)
    return u, kept
