"""ctypes driver of tests/_build/libpgmm_hostlogic.so (product host logic + reference-backed device stages, CPU only)."""
import ctypes as C
import os

from oracle import refmm2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    import build_hostlogic
    lib = C.CDLL(build_hostlogic.build())
    return refmm2.bind(lib, only=("mm_set_opt", "mm_check_opt", "mm_event_identity"))


def map_all(lib, seqs, names, preset="asm10", k=None, min_dp_max=90, threads=1):
    io, mo = refmm2.make_options(lib, preset, k, min_dp_max)  # OUR mm_set_opt / mm_check_opt
    seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
    names = [s if isinstance(s, bytes) else s.encode() for s in names]
    n = len(seqs)
    sa, na = (C.c_char_p * n)(*seqs), (C.c_char_p * n)(*names)
    n_regs = (C.c_int * n)()
    regs = (C.POINTER(refmm2.mm_reg1_t) * n)()
    lib.pgmm_hostlogic_map_all.restype = C.c_int
    rc = lib.pgmm_hostlogic_map_all(refmm2.REF_SO.encode(), n, sa, na, C.byref(io), C.byref(mo), threads, n_regs, regs)
    assert rc == 0
    out = []
    for i in range(n):
        out.append([refmm2.reg_to_tuple(lib, regs[i][j]) for j in range(n_regs[i])])
        for j in range(n_regs[i]):
            if regs[i][j].p:
                refmm2._libc.free(C.cast(regs[i][j].p, C.c_void_p))
        if regs[i]:
            refmm2._libc.free(C.cast(regs[i], C.c_void_p))
    return out, mo.mid_occ
