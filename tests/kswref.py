"""ctypes helpers to call the reference's ksw_extd2_sse and the plain-C oracle restatement on the same problem."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ksw_extz_t(C.Structure):  # C/ksw2.h:31-40
    _fields_ = [("max_zd", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int), ("mqe_t", C.c_int),
                ("mte", C.c_int), ("mte_q", C.c_int), ("score", C.c_int), ("m_cigar", C.c_int), ("n_cigar", C.c_int),
                ("reach_end", C.c_int), ("cigar", C.POINTER(C.c_uint32))]


class orc_ez_t(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score",
                                         "reach_end", "n_cigar")]


def simple_mat(a, b, sc_ambi):  # ksw_gen_simple_mat, C/align.c:9-22
    m = np.full((5, 5), -abs(b), dtype=np.int8)
    for i in range(4):
        m[i, i] = abs(a)
    m[4, :] = -abs(sc_ambi)
    m[:, 4] = -abs(sc_ambi)
    return np.ascontiguousarray(m.reshape(-1))


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def ref_extd2(lib, q, t, mat, gq, ge, gq2, ge2, w, zdrop, end_bonus, flag):
    ez = ksw_extz_t()
    lib.ksw_extd2_sse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int8, C.c_void_p,
                                  C.c_int8, C.c_int8, C.c_int8, C.c_int8, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(ksw_extz_t)]
    lib.ksw_extd2_sse.restype = None
    lib.ksw_extd2_sse(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, gq, ge, gq2, ge2, w,
                      zdrop, end_bonus, flag, C.byref(ez))
    cig = tuple(ez.cigar[i] for i in range(ez.n_cigar))
    if ez.cigar:
        _libc.free(C.cast(ez.cigar, C.c_void_p))
    return dict(max=ez.max_zd & 0x7fffffff, zdropped=ez.max_zd >> 31, max_q=ez.max_q, max_t=ez.max_t, mqe=ez.mqe,
                mqe_t=ez.mqe_t, mte=ez.mte, mte_q=ez.mte_q, score=ez.score, reach_end=ez.reach_end, cigar=cig)


def load_oracle():
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(so):
        raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle oracle`")
    return C.CDLL(so)


def orc_extd2(lib, q, t, mat, gq, ge, gq2, ge2, w, zdrop, end_bonus, flag):
    ez = orc_ez_t()
    cig = np.zeros(len(q) + len(t) + 2, dtype=np.uint32)
    lib.orc_ksw_extd2(len(q), C.c_void_p(q.ctypes.data), len(t), C.c_void_p(t.ctypes.data), C.c_void_p(mat.ctypes.data),
                      gq, ge, gq2, ge2, w, zdrop, end_bonus, flag, C.byref(ez), C.c_void_p(cig.ctypes.data))
    return dict(max=ez.max, zdropped=ez.zdropped, max_q=ez.max_q, max_t=ez.max_t, mqe=ez.mqe, mqe_t=ez.mqe_t,
                mte=ez.mte, mte_q=ez.mte_q, score=ez.score, reach_end=ez.reach_end,
                cigar=tuple(int(c) for c in cig[:ez.n_cigar]))


def random_pair(rng, qlen, tlen, div=0.05, indel=0.01, n_frac=0.0, big_indel=0):
    """A target and a diverged query (codes 0..4): substitutions, 1-bp indels, optional Ns and one long indel."""
    t = rng.integers(0, 4, size=tlen).astype(np.uint8)
    out = []
    i = 0
    cut = rng.integers(0, max(1, tlen)) if big_indel else -1
    while i < tlen and len(out) < qlen:
        if i == cut:
            if rng.random() < 0.5:
                i += big_indel
            else:
                out.extend(rng.integers(0, 4, size=big_indel).tolist())
            cut = -1
            continue
        r = rng.random()
        if r < indel / 2:
            i += 1
        elif r < indel:
            out.append(int(rng.integers(0, 4)))
        else:
            c = int(t[i])
            if rng.random() < div:
                c = (c + int(rng.integers(1, 4))) % 4
            out.append(c)
            i += 1
    while len(out) < qlen:
        out.append(int(rng.integers(0, 4)))
    q = np.array(out[:qlen], dtype=np.uint8)
    if n_frac > 0:
        q[rng.random(qlen) < n_frac] = 4
        t = t.copy()
        t[rng.random(tlen) < n_frac] = 4
    return q, t


# the only four flag combinations that occur on pangraph's path (C/align.c:714,755,759,798)
FLAG_LEFT_EXT = 0x40 | 0x02 | 0x80
FLAG_FILL1 = 0x08
FLAG_FILL2 = 0
FLAG_RIGHT_EXT = 0x40


def orc_extd2_unbanded(lib, q, t, mat, gq, ge, gq2, ge2, zdrop, end_bonus, flag):
    ez = orc_ez_t()
    cig = np.zeros(len(q) + len(t) + 2, dtype=np.uint32)
    ovf = C.c_int(0)
    lib.orc_ksw_extd2_unbanded(len(q), C.c_void_p(q.ctypes.data), len(t), C.c_void_p(t.ctypes.data), C.c_void_p(mat.ctypes.data),
                               gq, ge, gq2, ge2, zdrop, end_bonus, flag, C.byref(ez), C.c_void_p(cig.ctypes.data), C.byref(ovf))
    return dict(max=ez.max, zdropped=ez.zdropped, max_q=ez.max_q, max_t=ez.max_t, mqe=ez.mqe, mqe_t=ez.mqe_t,
                mte=ez.mte, mte_q=ez.mte_q, score=ez.score, reach_end=ez.reach_end,
                cigar=tuple(int(c) for c in cig[:ez.n_cigar])), ovf.value
