"""ctypes binding of oracle/_ref/libmm2ref.so -- the UNMODIFIED vendored minimap2 C of the reference.

TEST INFRASTRUCTURE ONLY.  Nothing under pangraph_b200/ may import this module; it exists so that
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can run the reference's own
code next to the CUDA path.

The call sequence of `ref_map_all` replays what the reference's Rust wrapper does
(packages/minimap2/src/options.rs:84-138, packages/minimap2/src/index.rs:17-55,
packages/pangraph/src/align/minimap2_lib/align_with_minimap2_lib.rs:29-85):
mm_set_opt(NULL) -> mm_set_opt(preset) -> init_opts -> mm_check_opt -> mm_idx_str -> mm_mapopt_update
-> mm_map per query in input order.
"""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libmm2ref.so")

# flag constants: packages/minimap2-sys/minimap2/minimap.h:10-47
MM_F_NO_DIAG, MM_F_NO_DUAL, MM_F_CIGAR, MM_F_OUT_CG = 0x1, 0x2, 0x4, 0x20
MM_F_NO_LJOIN, MM_F_ALL_CHAINS, MM_F_RMQ = 0x400, 0x800000, 0x80000000


class mm_idxopt_t(C.Structure):  # minimap.h:119-123
    _fields_ = [("k", C.c_short), ("w", C.c_short), ("flag", C.c_short), ("bucket_bits", C.c_short),
                ("mini_batch_size", C.c_int64), ("batch_size", C.c_uint64)]


class mm_mapopt_t(C.Structure):  # minimap.h:125-181
    _fields_ = [("flag", C.c_int64), ("seed", C.c_int), ("sdust_thres", C.c_int), ("max_qlen", C.c_int),
                ("bw", C.c_int), ("bw_long", C.c_int), ("max_gap", C.c_int), ("max_gap_ref", C.c_int),
                ("max_frag_len", C.c_int), ("max_chain_skip", C.c_int), ("max_chain_iter", C.c_int),
                ("min_cnt", C.c_int), ("min_chain_score", C.c_int), ("chain_gap_scale", C.c_float),
                ("chain_skip_scale", C.c_float), ("rmq_size_cap", C.c_int), ("rmq_inner_dist", C.c_int),
                ("rmq_rescue_size", C.c_int), ("rmq_rescue_ratio", C.c_float), ("mask_level", C.c_float),
                ("mask_len", C.c_int), ("pri_ratio", C.c_float), ("best_n", C.c_int), ("alt_drop", C.c_float),
                ("a", C.c_int), ("b", C.c_int), ("q", C.c_int), ("e", C.c_int), ("q2", C.c_int), ("e2", C.c_int),
                ("sc_ambi", C.c_int), ("noncan", C.c_int), ("junc_bonus", C.c_int), ("zdrop", C.c_int),
                ("zdrop_inv", C.c_int), ("end_bonus", C.c_int), ("min_dp_max", C.c_int), ("min_ksw_len", C.c_int),
                ("anchor_ext_len", C.c_int), ("anchor_ext_shift", C.c_int), ("max_clip_ratio", C.c_float),
                ("rank_min_len", C.c_int), ("rank_frac", C.c_float), ("pe_ori", C.c_int), ("pe_bonus", C.c_int),
                ("mid_occ_frac", C.c_float), ("q_occ_frac", C.c_float), ("min_mid_occ", C.c_int32),
                ("max_mid_occ", C.c_int32), ("mid_occ", C.c_int32), ("max_occ", C.c_int32),
                ("max_max_occ", C.c_int32), ("occ_dist", C.c_int32), ("mini_batch_size", C.c_int64),
                ("max_sw_mat", C.c_int64), ("cap_kalloc", C.c_int64), ("split_prefix", C.c_char_p)]


class mm_idx_seq_t(C.Structure):  # minimap.h:75-80
    _fields_ = [("name", C.c_char_p), ("offset", C.c_uint64), ("len", C.c_uint32), ("is_alt", C.c_uint32)]


class mm_idx_t(C.Structure):  # minimap.h:82-92
    _fields_ = [("b", C.c_int32), ("w", C.c_int32), ("k", C.c_int32), ("flag", C.c_int32), ("n_seq", C.c_uint32),
                ("index", C.c_int32), ("n_alt", C.c_int32), ("seq", C.POINTER(mm_idx_seq_t)),
                ("S", C.c_void_p), ("B", C.c_void_p), ("I", C.c_void_p), ("km", C.c_void_p), ("h", C.c_void_p)]


class mm_extra_t(C.Structure):  # minimap.h:95-101 (flexible cigar[] follows)
    _fields_ = [("capacity", C.c_uint32), ("dp_score", C.c_int32), ("dp_max", C.c_int32), ("dp_max2", C.c_int32),
                ("n_ambi_ts", C.c_uint32), ("n_cigar", C.c_uint32)]


class mm_reg1_t(C.Structure):  # minimap.h:103-117
    _fields_ = [("id", C.c_int32), ("cnt", C.c_int32), ("rid", C.c_int32), ("score", C.c_int32),
                ("qs", C.c_int32), ("qe", C.c_int32), ("rs", C.c_int32), ("re", C.c_int32),
                ("parent", C.c_int32), ("subsc", C.c_int32), ("as_", C.c_int32), ("mlen", C.c_int32),
                ("blen", C.c_int32), ("n_sub", C.c_int32), ("score0", C.c_int32), ("bits", C.c_uint32),
                ("hash", C.c_uint32), ("div", C.c_float), ("p", C.POINTER(mm_extra_t))]


assert C.sizeof(mm_idxopt_t) == 24 and C.sizeof(mm_mapopt_t) == 248
assert C.sizeof(mm_idx_t) == 80 and C.sizeof(mm_reg1_t) == 80 and C.sizeof(mm_extra_t) == 24

_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


_PROTOS = {
    "mm_set_opt": ([C.c_char_p, C.POINTER(mm_idxopt_t), C.POINTER(mm_mapopt_t)], C.c_int),
    "mm_check_opt": ([C.POINTER(mm_idxopt_t), C.POINTER(mm_mapopt_t)], C.c_int),
    "mm_idxopt_init": ([C.POINTER(mm_idxopt_t)], None),
    "mm_mapopt_init": ([C.POINTER(mm_mapopt_t)], None),
    "mm_idx_str": ([C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)],
                   C.POINTER(mm_idx_t)),
    "mm_mapopt_update": ([C.POINTER(mm_mapopt_t), C.POINTER(mm_idx_t)], None),
    "mm_idx_destroy": ([C.POINTER(mm_idx_t)], None),
    "mm_tbuf_init": ([], C.c_void_p),
    "mm_tbuf_destroy": ([C.c_void_p], None),
    "mm_map": ([C.POINTER(mm_idx_t), C.c_int, C.c_char_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(mm_mapopt_t),
                C.c_char_p], C.POINTER(mm_reg1_t)),
    "mm_event_identity": ([C.POINTER(mm_reg1_t)], C.c_double),
}


def bind(lib, only=None):
    """Attach prototypes of the boundary symbols to a loaded library (`only`: a subset, for the CPU-only
    host-logic test build, which exports just the option functions and mm_event_identity)."""
    for name, (argtypes, restype) in _PROTOS.items():
        if only is not None and name not in only:
            continue
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argtypes, restype
    return lib


def load_ref():
    if not os.path.exists(REF_SO):
        raise RuntimeError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference is mounted")
    return bind(C.CDLL(REF_SO))


def make_options(lib, preset="asm10", k=None, min_dp_max=90):
    """Minimap2Options::new + init_opts for pangraph's Minimap2Args{x,k,c,X,s,bucket_bits:14}."""
    io, mo = mm_idxopt_t(), mm_mapopt_t()
    if lib.mm_set_opt(None, C.byref(io), C.byref(mo)) != 0:
        raise RuntimeError("mm_set_opt(NULL) failed")
    if lib.mm_set_opt(preset.encode(), C.byref(io), C.byref(mo)) != 0:
        raise ValueError(f"unknown preset {preset}")
    if k is not None:
        io.k = k
    mo.flag |= MM_F_OUT_CG | MM_F_CIGAR
    mo.min_dp_max = min_dp_max
    mo.flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN
    io.bucket_bits = 14
    rc = lib.mm_check_opt(C.byref(io), C.byref(mo))
    if rc != 0:
        raise ValueError(f"mm_check_opt -> {rc}")
    return io, mo


def reg_to_tuple(lib, r):
    """Flatten one mm_reg1_t (+ its mm_extra_t and CIGAR) to a plain comparable tuple; every field."""
    base = (r.id, r.cnt, r.rid, r.score, r.qs, r.qe, r.rs, r.re, r.parent, r.subsc, r.as_, r.mlen, r.blen,
            r.n_sub, r.score0, r.bits, r.hash, C.c_uint32.from_buffer_copy(C.c_float(r.div)).value)
    if not r.p:
        return base + (None,)
    p = r.p.contents
    cig_t = C.c_uint32 * p.n_cigar
    cig = tuple(cig_t.from_address(C.addressof(p) + 24))
    de = 1.0 - lib.mm_event_identity(C.byref(r))
    return base + ((p.capacity, p.dp_score, p.dp_max, p.dp_max2, p.n_ambi_ts, cig, de),)


def cigar_str(cig):
    return "".join(f"{c >> 4}{'MIDNSHP=XB'[c & 0xf]}" for c in cig)


class Index:
    def __init__(self, lib, seqs, names, preset="asm10", k=None, min_dp_max=90):
        self.lib = lib
        self.io, self.mo = make_options(lib, preset, k, min_dp_max)
        self.seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
        self.names = [s if isinstance(s, bytes) else s.encode() for s in names]
        n = len(self.seqs)
        sa = (C.c_char_p * n)(*self.seqs)
        na = (C.c_char_p * n)(*self.names)
        self.mi = lib.mm_idx_str(self.io.w, self.io.k, self.io.flag & 1, self.io.bucket_bits, n, sa, na)
        if not self.mi:
            raise RuntimeError("minimap2: failed to create index")
        lib.mm_mapopt_update(C.byref(self.mo), self.mi)

    def map_one(self, i):
        lib = self.lib
        tb = lib.mm_tbuf_init()
        n = C.c_int(0)
        regs = lib.mm_map(self.mi, len(self.seqs[i]), self.seqs[i], C.byref(n), tb, C.byref(self.mo), self.names[i])
        out = [reg_to_tuple(lib, regs[j]) for j in range(n.value)]
        for j in range(n.value):
            if regs[j].p:
                _libc.free(C.cast(regs[j].p, C.c_void_p))
        if regs:
            _libc.free(C.cast(regs, C.c_void_p))
        lib.mm_tbuf_destroy(tb)
        return out

    def map_all(self, threads=1):
        idx = range(len(self.seqs))
        if threads <= 1:
            return [self.map_one(i) for i in idx]
        with ThreadPoolExecutor(threads) as ex:  # ctypes drops the GIL inside mm_map
            return list(ex.map(self.map_one, idx))

    def close(self):
        if self.mi:
            self.lib.mm_idx_destroy(self.mi)
            self.mi = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_map_all(seqs, names, preset="asm10", k=None, min_dp_max=90, threads=1):
    idx = Index(load_ref(), seqs, names, preset, k, min_dp_max)
    try:
        return idx.map_all(threads), idx.mo.mid_occ
    finally:
        idx.close()


# ---- stage-level access to the reference (for the K1 / K3 parity tests) ----

class mm128_t(C.Structure):
    _fields_ = [("x", C.c_uint64), ("y", C.c_uint64)]


class mm128_v(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(mm128_t))]


def ref_sketch(lib, seq, w, k, rid=0):
    """mm_sketch (sketch.c:77-143) -> list of (x, y)."""
    seq = seq if isinstance(seq, bytes) else seq.encode()
    v = mm128_v(0, 0, None)
    lib.mm_sketch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(mm128_v)]
    lib.mm_sketch.restype = None
    if len(seq) == 0:
        return []
    lib.mm_sketch(None, seq, len(seq), w, k, rid, 0, C.byref(v))
    out = [(v.a[i].x, v.a[i].y) for i in range(v.n)]
    if v.a:
        _libc.free(C.cast(v.a, C.c_void_p))
    return out


# ---- the seeding stage of the reference on its own (for the K1+K3 parity test) ----
STAGE_SO = os.path.join(HERE, "_ref", "libmm2ref_stage.so")


def load_ref_stage():
    """oracle/_ref/libmm2ref_stage.so: the same unmodified sources with map.c's file-local functions visible."""
    if not os.path.exists(STAGE_SO):
        raise RuntimeError(f"{STAGE_SO} missing: run `make -C oracle ref` where /root/reference is mounted")
    return bind(C.CDLL(STAGE_SO))


def ref_collect_seed_hits(lib, idx, seq, name):
    """What mm_map_frag does up to the chaining call (map.c:250-253): collect_minimizers, mm_seed_mz_flt,
    collect_seed_hits.  -> (anchors [(x, y)] as the reference's radix sort leaves them, mini_pos [u64], rep_len)."""
    seq = seq if isinstance(seq, bytes) else seq.encode()
    name = name if isinstance(name, bytes) else name.encode()
    mv = mm128_v(0, 0, None)
    qlen = C.c_int(len(seq))
    sp = C.c_char_p(seq)
    lib.collect_minimizers.argtypes = [C.c_void_p, C.POINTER(mm_mapopt_t), C.POINTER(mm_idx_t), C.c_int, C.POINTER(C.c_int),
                                       C.POINTER(C.c_char_p), C.POINTER(mm128_v)]
    lib.collect_minimizers.restype = None
    lib.collect_minimizers(None, C.byref(idx.mo), idx.mi, 1, C.byref(qlen), C.byref(sp), C.byref(mv))
    if idx.mo.q_occ_frac > 0.0:
        lib.mm_seed_mz_flt.argtypes = [C.c_void_p, C.POINTER(mm128_v), C.c_int32, C.c_float]
        lib.mm_seed_mz_flt.restype = None
        lib.mm_seed_mz_flt(None, C.byref(mv), idx.mo.mid_occ, idx.mo.q_occ_frac)
    n_a, rep_len, n_mini = C.c_int64(0), C.c_int(0), C.c_int(0)
    mini = C.POINTER(C.c_uint64)()
    lib.collect_seed_hits.argtypes = [C.c_void_p, C.POINTER(mm_mapopt_t), C.c_int, C.POINTER(mm_idx_t), C.c_char_p, C.POINTER(mm128_v),
                                      C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.POINTER(C.c_uint64))]
    lib.collect_seed_hits.restype = C.POINTER(mm128_t)
    a = lib.collect_seed_hits(None, C.byref(idx.mo), idx.mo.mid_occ, idx.mi, name, C.byref(mv), len(seq), C.byref(n_a),
                              C.byref(rep_len), C.byref(n_mini), C.byref(mini))
    import numpy as np
    anchors = np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_uint64)), shape=(n_a.value, 2)).copy() if n_a.value else np.zeros((0, 2), np.uint64)
    mp = np.ctypeslib.as_array(mini, shape=(n_mini.value,)).copy() if n_mini.value else np.zeros(0, np.uint64)
    for ptr in (a, mini, mv.a):
        if ptr:
            _libc.free(C.cast(ptr, C.c_void_p))
    return anchors, mp, rep_len.value
