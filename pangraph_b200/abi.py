"""ctypes binding of libpgmm_b200.so (include/pgmm_b200.h).  Plumbing only: every call lands in the CUDA library;
there is no Python or CPU implementation behind these functions, and loading fails loudly if the library is absent."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpgmm_b200.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "pangraph_b200 has no fallback implementation")
        _lib = C.CDLL(LIB_PATH)
        _lib.pgmm_device_count.restype = C.c_int
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def ksw_extd2_batch(qlen, tlen, q_off, t_off, qcodes, tcodes, w, zdrop, end_bonus, flag, a, b, sc_ambi, q, e, q2, e2,
                    arena_budget_bytes=0):
    """Stage K5 on host buffers. Returns (ez[n,11] int32, list of cigar tuples, kernel_ms)."""
    L = lib()
    n = len(qlen)
    qlen = np.ascontiguousarray(qlen, dtype=np.int32)
    tlen = np.ascontiguousarray(tlen, dtype=np.int32)
    q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
    t_off = np.ascontiguousarray(t_off, dtype=np.uint64)
    w = np.ascontiguousarray(w, dtype=np.int32)
    zdrop = np.ascontiguousarray(zdrop, dtype=np.int32)
    end_bonus = np.ascontiguousarray(end_bonus, dtype=np.int32)
    flag = np.ascontiguousarray(flag, dtype=np.int32)
    qcodes = np.ascontiguousarray(qcodes, dtype=np.uint8)
    tcodes = np.ascontiguousarray(tcodes, dtype=np.uint8)
    ez = np.zeros((n, 11), dtype=np.int32)
    cap = int((qlen.astype(np.int64) + tlen + 2).sum()) + 16
    cig = np.zeros(cap, dtype=np.uint32)
    start = np.zeros(n, dtype=np.uint64)
    ms = C.c_double(0)
    rc = L.pgmm_ksw_extd2_batch(C.c_int(n), _p(qlen), _p(tlen), _p(q_off), _p(t_off), _p(qcodes), C.c_uint64(qcodes.size),
                                _p(tcodes), C.c_uint64(tcodes.size), _p(w), _p(zdrop), _p(end_bonus), _p(flag),
                                C.c_int(a), C.c_int(b), C.c_int(sc_ambi), C.c_int(q), C.c_int(e), C.c_int(q2), C.c_int(e2),
                                _p(ez), _p(cig), C.c_uint64(cap), _p(start), C.byref(ms), C.c_uint64(arena_budget_bytes))
    if rc != 0:
        raise RuntimeError(f"pgmm_ksw_extd2_batch -> {rc}")
    cigs = [tuple(int(c) for c in cig[int(start[i]):int(start[i]) + int(ez[i, 10])]) for i in range(n)]
    return ez, cigs, ms.value
