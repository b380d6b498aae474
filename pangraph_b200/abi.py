"""ctypes binding of libpgmm_b200.so (include/pgmm_b200.h).  Plumbing only: every call lands in the CUDA library;
there is no Python or CPU implementation behind these functions, and loading fails loudly if the library is absent."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpgmm_b200.so")

# flag constants: include/pgmm_b200.h (reference: minimap2/minimap.h:10-47)
MM_F_NO_DIAG, MM_F_NO_DUAL, MM_F_CIGAR, MM_F_OUT_CG = 0x1, 0x2, 0x4, 0x20
MM_F_NO_LJOIN, MM_F_ALL_CHAINS, MM_F_RMQ = 0x400, 0x800000, 0x80000000


class mm_idxopt_t(C.Structure):
    _fields_ = [("k", C.c_short), ("w", C.c_short), ("flag", C.c_short), ("bucket_bits", C.c_short),
                ("mini_batch_size", C.c_int64), ("batch_size", C.c_uint64)]


class mm_mapopt_t(C.Structure):
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


class mm_idx_seq_t(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_uint64), ("len", C.c_uint32), ("is_alt", C.c_uint32)]


class mm_idx_t(C.Structure):
    _fields_ = [("b", C.c_int32), ("w", C.c_int32), ("k", C.c_int32), ("flag", C.c_int32), ("n_seq", C.c_uint32),
                ("index", C.c_int32), ("n_alt", C.c_int32), ("seq", C.POINTER(mm_idx_seq_t)),
                ("S", C.c_void_p), ("B", C.c_void_p), ("I", C.c_void_p), ("km", C.c_void_p), ("h", C.c_void_p)]


class mm_extra_t(C.Structure):
    _fields_ = [("capacity", C.c_uint32), ("dp_score", C.c_int32), ("dp_max", C.c_int32), ("dp_max2", C.c_int32),
                ("n_ambi_ts", C.c_uint32), ("n_cigar", C.c_uint32)]


class mm_reg1_t(C.Structure):
    _fields_ = [("id", C.c_int32), ("cnt", C.c_int32), ("rid", C.c_int32), ("score", C.c_int32),
                ("qs", C.c_int32), ("qe", C.c_int32), ("rs", C.c_int32), ("re", C.c_int32),
                ("parent", C.c_int32), ("subsc", C.c_int32), ("as_", C.c_int32), ("mlen", C.c_int32),
                ("blen", C.c_int32), ("n_sub", C.c_int32), ("score0", C.c_int32), ("bits", C.c_uint32),
                ("hash", C.c_uint32), ("div", C.c_float), ("p", C.POINTER(mm_extra_t))]


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]

_lib = None


def lib():
    global _lib
    if _lib is None:
        # one stream per round in flight and per DP size class: more hardware queues than the default 8 (only
        # effective if CUDA has not been initialised in this process yet)
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "pangraph_b200 has no fallback implementation")
        _lib = C.CDLL(LIB_PATH)
        _lib.pgmm_device_count.restype = C.c_int
        _lib.mm_set_opt.argtypes = [C.c_char_p, C.POINTER(mm_idxopt_t), C.POINTER(mm_mapopt_t)]
        _lib.mm_check_opt.argtypes = [C.POINTER(mm_idxopt_t), C.POINTER(mm_mapopt_t)]
        _lib.mm_idx_str.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        _lib.mm_idx_str.restype = C.POINTER(mm_idx_t)
        _lib.mm_mapopt_update.argtypes = [C.POINTER(mm_mapopt_t), C.POINTER(mm_idx_t)]
        _lib.mm_mapopt_update.restype = None
        _lib.mm_idx_destroy.argtypes = [C.POINTER(mm_idx_t)]
        _lib.mm_idx_destroy.restype = None
        _lib.mm_tbuf_init.restype = C.c_void_p
        _lib.mm_tbuf_destroy.argtypes = [C.c_void_p]
        _lib.mm_tbuf_destroy.restype = None
        _lib.mm_map.argtypes = [C.POINTER(mm_idx_t), C.c_int, C.c_char_p, C.POINTER(C.c_int), C.c_void_p,
                                C.POINTER(mm_mapopt_t), C.c_char_p]
        _lib.mm_map.restype = C.POINTER(mm_reg1_t)
        _lib.mm_event_identity.argtypes = [C.POINTER(mm_reg1_t)]
        _lib.mm_event_identity.restype = C.c_double
        _lib.pgmm_map_batch.restype = None
        _lib.pgmm_map_self.restype = None
        _lib.pgmm_idx_upload.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        _lib.pgmm_idx_upload.restype = C.POINTER(mm_idx_t)
        _lib.pgmm_idx_build.argtypes = [C.POINTER(mm_idx_t), C.c_int, C.c_int, C.c_int]
        _lib.pgmm_idx_build.restype = None
        _lib.pgmm_get_stats.restype = None
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


def chain_rmq(a, max_dist, max_dist_inner, bw, max_skip, cap, min_cnt, min_sc, pen_gap, pen_skip, host_redo=True):
    """Stage K4 (+ host backtrack) on sorted anchors a[n,2] uint64.  Returns (u, kept anchors, fpv[3,n] int32,
    (segments, handed back, their anchors)); raises if the device handed segments back and host_redo is False."""
    L = lib()
    L.pgmm_chain_rmq.restype = C.c_int64
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    n = len(a)
    u = np.zeros(n + 1, dtype=np.uint64)
    fpv = np.zeros((3, max(n, 1)), dtype=np.int32)
    seg = np.zeros(3, dtype=np.int64)
    n_a = C.c_int64(0)
    n_u = L.pgmm_chain_rmq(_p(a), C.c_int64(n), C.c_int(max_dist), C.c_int(max_dist_inner), C.c_int(bw), C.c_int(max_skip),
                           C.c_int(cap), C.c_int(min_cnt), C.c_int(min_sc), C.c_float(pen_gap), C.c_float(pen_skip), _p(u),
                           C.byref(n_a), _p(fpv), _p(seg), C.c_int(1 if host_redo else 0))
    if n_u < 0:
        raise RuntimeError(f"pgmm_chain_rmq -> {n_u} ({int(seg[1])} of {int(seg[0])} segments handed back)")
    return u[:n_u].copy(), a[:n_a.value].copy(), fpv[:, :n], tuple(int(v) for v in seg)


def set_device(device):
    rc = lib().pgmm_set_device(int(device))
    if rc != 0:
        raise RuntimeError(f"pgmm_set_device({device}) -> {rc}")


STAT_NAMES = ("total_ms", "seed_ms", "dp_kernel_ms", "index_ms", "dp_jobs", "dp_cells", "dp_waves", "bases_mapped",
              "bases_indexed", "batches", "launches", "t_encode", "t_seed", "t_chain", "t_dp", "t_stitch", "t_final",
              "h2d_bytes", "d2h_bytes", "dp_seq_bytes", "device_mallocs",
              "k5_ms", "k5_cells", "k5_bases", "k5_launches", "k5a_ms", "k5a_cells", "k5a_bases", "k5a_launches",
              "k5b_ms", "k5b_cells", "k5b_bases", "k5b_launches",
              "t_chain_sort", "t_chain_fill", "t_chain_rest", "chain_kernel_ms", "chain_anchors", "chain_segments",
              "chain_redo_segments", "chain_redo_anchors", "chain_launches", "chain_iterations", "chain_batches",
              "cpu_encode", "cpu_seed", "cpu_sort", "cpu_chain_fill", "cpu_backtrack_plan", "cpu_dp_round_side", "cpu_dp_workers",
              "cpu_stitch", "cpu_final", "cpu_index", "anchor_sort_device", "anchor_sort_host")


def get_stats(reset=False):
    out = (C.c_double * len(STAT_NAMES))()
    lib().pgmm_get_stats(out, len(STAT_NAMES), 1 if reset else 0)
    return dict(zip(STAT_NAMES, out))


def make_options(preset="asm10", k=None, min_dp_max=90):
    """What the reference's wrapper builds for pangraph: Minimap2Options::new + init_opts with
    Minimap2Args{x:preset, k, c:true, X:true, s:min_dp_max, bucket_bits:14}
    (packages/minimap2/src/options.rs:84-138, options_args.rs:273-331; align_with_minimap2_lib.rs:49-57)."""
    L = lib()
    io, mo = mm_idxopt_t(), mm_mapopt_t()
    if L.mm_set_opt(None, C.byref(io), C.byref(mo)) != 0:
        raise RuntimeError("mm_set_opt(NULL) failed")
    if L.mm_set_opt(preset.encode(), C.byref(io), C.byref(mo)) != 0:
        raise ValueError(f"minimap2: mm_set_opt(preset, ...): failed to set options: incorrect preset {preset}")
    if k is not None:
        io.k = k
    mo.flag |= MM_F_OUT_CG | MM_F_CIGAR
    mo.min_dp_max = min_dp_max
    mo.flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN
    io.bucket_bits = 14
    rc = L.mm_check_opt(C.byref(io), C.byref(mo))
    if rc != 0:
        raise ValueError(f"minimap2: mm_check_opt(): options are invalid ({rc})")
    return io, mo


def reg_to_tuple(r):
    """One mm_reg1_t (+ mm_extra_t + CIGAR + de) as a plain tuple; same field order as oracle.refmm2.reg_to_tuple."""
    base = (r.id, r.cnt, r.rid, r.score, r.qs, r.qe, r.rs, r.re, r.parent, r.subsc, r.as_, r.mlen, r.blen,
            r.n_sub, r.score0, r.bits, r.hash, C.c_uint32.from_buffer_copy(C.c_float(r.div)).value)
    if not r.p:
        return base + (None,)
    p = r.p.contents
    cig = tuple((C.c_uint32 * p.n_cigar).from_address(C.addressof(p) + 24))
    de = 1.0 - lib().mm_event_identity(C.byref(r))
    return base + ((p.capacity, p.dp_score, p.dp_max, p.dp_max2, p.n_ambi_ts, cig, de),)


def _take_regs(regs, n):
    out = [reg_to_tuple(regs[j]) for j in range(n)]
    for j in range(n):
        if regs[j].p:
            _libc.free(C.cast(regs[j].p, C.c_void_p))
    if regs:
        _libc.free(C.cast(regs, C.c_void_p))
    return out


class Index:
    """Minimap2Index of the reference wrapper (packages/minimap2/src/index.rs:17-55) over the GPU library."""

    def __init__(self, seqs, names, preset="asm10", k=None, min_dp_max=90, resident_only=False):
        """resident_only=True stops after pgmm_idx_upload (bases coded and copied to HBM); call build() next."""
        L = lib()
        self.io, self.mo = make_options(preset, k, min_dp_max)
        self.seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
        self.names = [s if isinstance(s, bytes) else s.encode() for s in names]
        n = len(self.seqs)
        sa = (C.c_char_p * n)(*self.seqs)
        na = (C.c_char_p * n)(*self.names)
        if resident_only:
            self.mi = L.pgmm_idx_upload(n, sa, na)
            if not self.mi:
                raise RuntimeError("minimap2: failed to create index")
            return
        self.mi = L.mm_idx_str(self.io.w, self.io.k, self.io.flag & 1, self.io.bucket_bits, n, sa, na)
        if not self.mi:
            raise RuntimeError("minimap2: failed to create index")
        L.mm_mapopt_update(C.byref(self.mo), self.mi)

    def build(self):
        """K1+K2 on the resident bases, then mm_mapopt_update."""
        L = lib()
        L.pgmm_idx_build(self.mi, self.io.w, self.io.k, self.io.bucket_bits)
        self.mo.mid_occ = 0
        L.mm_mapopt_update(C.byref(self.mo), self.mi)

    def map_self(self, raw=False):
        """pgmm_map_self: the indexed sequences against their own index, nothing re-uploaded."""
        L = lib()
        n = len(self.seqs)
        n_regs = (C.c_int * n)()
        regs = (C.POINTER(mm_reg1_t) * n)()
        L.pgmm_map_self(self.mi, C.byref(self.mo), n_regs, regs)
        if raw:
            return n_regs, regs
        return [_take_regs(regs[i], n_regs[i]) for i in range(n)]

    def map_one(self, seq, name):
        """Minimap2Mapper::run_map (packages/minimap2/src/map.rs:26-41): one mm_map call."""
        L = lib()
        seq = seq if isinstance(seq, bytes) else seq.encode()
        name = name if isinstance(name, bytes) else name.encode()
        tb = L.mm_tbuf_init()
        n = C.c_int(0)
        regs = L.mm_map(self.mi, len(seq), seq, C.byref(n), tb, C.byref(self.mo), name)
        out = _take_regs(regs, n.value)
        L.mm_tbuf_destroy(tb)
        return out

    def map_batch(self, seqs=None, names=None):
        """All queries of one round through pgmm_map_batch (default: the indexed sequences themselves)."""
        L = lib()
        seqs = self.seqs if seqs is None else [s if isinstance(s, bytes) else s.encode() for s in seqs]
        names = self.names if names is None else [s if isinstance(s, bytes) else s.encode() for s in names]
        n = len(seqs)
        sa, na = (C.c_char_p * n)(*seqs), (C.c_char_p * n)(*names)
        lens = (C.c_int * n)(*[len(s) for s in seqs])
        n_regs = (C.c_int * n)()
        regs = (C.POINTER(mm_reg1_t) * n)()
        L.pgmm_map_batch(self.mi, n, lens, sa, na, C.byref(self.mo), n_regs, regs)
        return [_take_regs(regs[i], n_regs[i]) for i in range(n)]

    def collect_seeds(self, seqs=None, names=None):
        """Stage K1+K3: per query (anchors[n,2] uint64 before the sort, mini_pos uint64[], rep_len)."""
        L = lib()
        seqs = self.seqs if seqs is None else [s if isinstance(s, bytes) else s.encode() for s in seqs]
        names = self.names if names is None else [s if isinstance(s, bytes) else s.encode() for s in names]
        n = len(seqs)
        sa, na = (C.c_char_p * n)(*seqs), (C.c_char_p * n)(*names)
        lens = (C.c_int * n)(*[len(s) for s in seqs])
        cap_a = 1 << 22
        while True:
            anchors = np.zeros((cap_a, 2), dtype=np.uint64)
            mini = np.zeros(sum(len(s) for s in seqs) + 16, dtype=np.uint64)
            counts = np.zeros((n, 3), dtype=np.int64)
            rc = L.pgmm_collect_seeds(self.mi, n, lens, sa, na, C.byref(self.mo), _p(anchors), C.c_uint64(cap_a), _p(mini),
                                      C.c_uint64(mini.size), _p(counts))
            if rc == 0:
                break
            cap_a *= 4
        out, a0, m0 = [], 0, 0
        for i in range(n):
            na_, nm_, rep = (int(x) for x in counts[i])
            out.append((anchors[a0:a0 + na_].copy(), mini[m0:m0 + nm_].copy(), rep))
            a0 += na_
            m0 += nm_
        return out

    def close(self):
        if self.mi:
            lib().mm_idx_destroy(self.mi)
            self.mi = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sketch(seqs, w, k):
    """Stage K1: list of (x[], y[]) uint64 arrays, one per sequence."""
    L = lib()
    seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
    n = len(seqs)
    sa = (C.c_char_p * n)(*seqs)
    lens = (C.c_int * n)(*[len(s) for s in seqs])
    cap = sum(len(s) for s in seqs) + 16
    x = np.zeros(cap, dtype=np.uint64)
    y = np.zeros(cap, dtype=np.uint64)
    off = np.zeros(n + 1, dtype=np.uint64)
    rc = L.pgmm_sketch(n, sa, lens, w, k, _p(x), _p(y), C.c_uint64(cap), _p(off))
    if rc != 0:
        raise RuntimeError(f"pgmm_sketch -> {rc}")
    return [(x[int(off[i]):int(off[i + 1])].copy(), y[int(off[i]):int(off[i + 1])].copy()) for i in range(n)]


# ---------------- host half of the path (include/pgmm_b200.h part 2) ----------------

class pgmm_alignment_t(C.Structure):
    _fields_ = [("qry_name", C.c_uint64), ("ref_name", C.c_uint64), ("qry_len", C.c_uint64), ("ref_len", C.c_uint64),
                ("qry_start", C.c_uint64), ("qry_end", C.c_uint64), ("ref_start", C.c_uint64), ("ref_end", C.c_uint64),
                ("matches", C.c_uint64), ("length", C.c_uint64), ("quality", C.c_uint64), ("reverse", C.c_int32),
                ("has_divergence", C.c_int32), ("divergence", C.c_double), ("align", C.c_double), ("n_cigar", C.c_uint32),
                ("cigar", C.POINTER(C.c_uint32))]


class pgmm_alignment_args_t(C.Structure):
    _fields_ = [("indel_len_threshold", C.c_uint64), ("alpha", C.c_double), ("beta", C.c_double),
                ("sensitivity", C.c_uint64), ("kmer_length", C.c_int64)]


_OPS = "MIDNSHP=XB"


def alignment_args(indel_len_threshold=100, alpha=100.0, beta=10.0, sensitivity=10, kmer_length=None):
    """AlignmentArgs of the reference (packages/pangraph/src/align/alignment_args.rs) with its defaults."""
    return pgmm_alignment_args_t(indel_len_threshold, alpha, beta, sensitivity, kmer_length or 0)


def _aln_to_c(a, keep):
    cig = (C.c_uint32 * max(1, len(a["cigar"])))(*[(n << 4) | _OPS.index(op) for n, op in a["cigar"]])
    keep.append(cig)
    d = a.get("divergence")
    return pgmm_alignment_t(a["qry"][0], a["ref"][0], a["qry"][1], a["ref"][1], a["qry"][2], a["qry"][3], a["ref"][2], a["ref"][3],
                            a["matches"], a["length"], a["quality"], 1 if a["reverse"] else 0, 0 if d is None else 1,
                            0.0 if d is None else d, a.get("align") or 0.0, len(a["cigar"]), C.cast(cig, C.POINTER(C.c_uint32)))


def _aln_from_c(c):
    return dict(qry=(c.qry_name, c.qry_len, c.qry_start, c.qry_end), ref=(c.ref_name, c.ref_len, c.ref_start, c.ref_end),
                matches=c.matches, length=c.length, quality=c.quality, reverse=bool(c.reverse),
                cigar=[(c.cigar[i] >> 4, _OPS[c.cigar[i] & 15]) for i in range(c.n_cigar)],
                divergence=c.divergence if c.has_divergence else None, align=c.align)


def _take_alns(out, n):
    res = [_aln_from_c(out[i]) for i in range(n.value)]
    lib().pgmm_alignments_free(out, n)
    return res


def split_matches(aln, args):
    keep = []
    c = _aln_to_c(aln, keep)
    out, n = C.POINTER(pgmm_alignment_t)(), C.c_size_t(0)
    rc = lib().pgmm_split_matches(C.byref(c), C.byref(args), C.byref(out), C.byref(n))
    if rc != 0:
        raise ValueError(f"split_matches: unexpected CIGAR operation ({rc})")
    return _take_alns(out, n)


def alignment_energy2(aln, args):
    keep = []
    c = _aln_to_c(aln, keep)
    lib().pgmm_alignment_energy2.restype = C.c_double
    return lib().pgmm_alignment_energy2(C.byref(c), C.byref(args))


def filter_matches(alns, args):
    keep = []
    arr = (pgmm_alignment_t * max(1, len(alns)))(*[_aln_to_c(a, keep) for a in alns])
    out, n = C.POINTER(pgmm_alignment_t)(), C.c_size_t(0)
    lib().pgmm_filter_matches(arr, C.c_size_t(len(alns)), C.byref(args), C.byref(out), C.byref(n))
    return _take_alns(out, n)


def _blocks(blocks):
    ids = list(blocks.keys())
    n = len(ids)
    seqs = [blocks[i] if isinstance(blocks[i], bytes) else blocks[i].encode() for i in ids]
    return n, (C.c_uint64 * max(1, n))(*ids), (C.c_char_p * max(1, n))(*seqs), seqs


def align_with_minimap2_lib(blocks, args):
    """blocks: {block_id: consensus}; returns the Alignment list of one round
    (packages/pangraph/src/align/minimap2_lib/align_with_minimap2_lib.rs:15-27)."""
    n, ids, sa, keep = _blocks(blocks)
    out, cnt = C.POINTER(pgmm_alignment_t)(), C.c_size_t(0)
    rc = lib().pgmm_align_with_minimap2_lib(n, ids, sa, C.byref(args), C.byref(out), C.byref(cnt))
    if rc == -1:
        raise ValueError(f"Unknown sensitivity preset: {args.sensitivity}")
    if rc != 0:
        raise RuntimeError(f"align_with_minimap2_lib -> {rc}")
    return _take_alns(out, cnt)


def find_filtered_matches(blocks, args):
    """find_matches + self-hit removal + split_matches + filter_matches (graph_merging.rs:95-121)."""
    n, ids, sa, keep = _blocks(blocks)
    out, cnt = C.POINTER(pgmm_alignment_t)(), C.c_size_t(0)
    rc = lib().pgmm_find_filtered_matches(n, ids, sa, C.byref(args), C.byref(out), C.byref(cnt))
    if rc == -1:
        raise ValueError(f"Unknown sensitivity preset: {args.sensitivity}")
    if rc != 0:
        raise RuntimeError(f"find_filtered_matches -> {rc}")
    return _take_alns(out, cnt)


# ---------------- map_variations (include/pgmm_b200.h part 4) ----------------

class pgmm_edit_t(C.Structure):
    _fields_ = [("status", C.c_int32), ("hit_boundary", C.c_int32), ("attempts", C.c_int32), ("band_width", C.c_int32),
                ("score", C.c_int32), ("n_sub", C.c_int32), ("n_del", C.c_int32), ("n_ins", C.c_int32),
                ("sub_pos", C.POINTER(C.c_int32)), ("sub_chr", C.POINTER(C.c_char)), ("del_pos", C.POINTER(C.c_int32)),
                ("del_len", C.POINTER(C.c_int32)), ("ins_pos", C.POINTER(C.c_int32)), ("ins_len", C.POINTER(C.c_int32)),
                ("ins_seq", C.POINTER(C.c_char))]


def map_variations_batch(refs, qrys, mean_shifts, band_widths, extra_band_width=5, max_alignment_attempts=4, with_stats=False):
    """map_variations (packages/pangraph/src/align/map_variations.rs:39-80) for a batch of (ref, qry) pairs.
    -> per problem dict(subs=[(pos, chr)], dels=[(pos, len)], inss=[(pos, seq)], hit_boundary, attempts, score) or the negative
    status where the reference returns Err / panics."""
    L = lib()
    n = len(refs)
    refs = [s if isinstance(s, bytes) else s.encode() for s in refs]
    qrys = [s if isinstance(s, bytes) else s.encode() for s in qrys]
    ra, qa = (C.c_char_p * n)(*refs), (C.c_char_p * n)(*qrys)
    rl, ql = (C.c_int32 * n)(*[len(s) for s in refs]), (C.c_int32 * n)(*[len(s) for s in qrys])
    ms, bw = (C.c_int32 * n)(*mean_shifts), (C.c_int32 * n)(*band_widths)
    out = C.POINTER(pgmm_edit_t)()
    st = (C.c_double * 4)()
    L.pgmm_map_variations_batch.restype = C.c_int
    rc = L.pgmm_map_variations_batch(n, ra, rl, qa, ql, ms, bw, extra_band_width, max_alignment_attempts, C.byref(out), st)
    if rc != 0:
        raise RuntimeError(f"pgmm_map_variations_batch -> {rc}")
    res = []
    for i in range(n):
        e = out[i]
        if e.status != 0:
            res.append(int(e.status))
            continue
        ins, off = [], 0
        for k in range(e.n_ins):
            ins.append((int(e.ins_pos[k]), C.string_at(C.addressof(e.ins_seq.contents) + off, e.ins_len[k]).decode()))
            off += e.ins_len[k]
        res.append(dict(subs=[(int(e.sub_pos[k]), e.sub_chr[k].decode()) for k in range(e.n_sub)],
                        dels=[(int(e.del_pos[k]), int(e.del_len[k])) for k in range(e.n_del)], inss=ins,
                        hit_boundary=bool(e.hit_boundary), attempts=int(e.attempts), score=int(e.score)))
    L.pgmm_edits_free.argtypes = [C.POINTER(pgmm_edit_t), C.c_int]
    L.pgmm_edits_free(out, n)
    if with_stats:
        return res, dict(kernel_ms=st[0], cells=st[1], problems=st[2], launches=st[3])
    return res
