"""ctypes helpers for the nextalign / map_variations restatement in oracle/liboracle.so (oracle/nextalign_oracle.c)."""
import ctypes as C

import numpy as np

import kswref


class orc_na_params_t(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("penalty_gap_extend", "penalty_gap_open", "penalty_mismatch", "score_match",
                                         "left_terminal_gaps_free", "right_terminal_gaps_free", "left_align", "min_length",
                                         "max_alignment_attempts")]


def params(min_length=1, max_alignment_attempts=3, **kw):
    lib = kswref.load_oracle()
    p = orc_na_params_t()
    lib.orc_na_default_params(C.byref(p), max_alignment_attempts)
    p.min_length = min_length
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def align_nuc_simplestripe(qry, ref, mean_shift, band_width, p):
    """-> dict(qry_aln, ref_aln, score, hit_boundary, band_width, attempts) or an int error code"""
    lib = kswref.load_oracle()
    lib.orc_align_nuc_simplestripe.restype = C.c_int64
    q, r = qry.encode(), ref.encode()
    cap = len(q) + len(r) + 2
    aq, ar = C.create_string_buffer(cap), C.create_string_buffer(cap)
    score, hb, bw, att = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    n = lib.orc_align_nuc_simplestripe(q, len(q), r, len(r), mean_shift, band_width, C.byref(p), aq, ar, C.byref(score), C.byref(hb),
                                       C.byref(bw), C.byref(att))
    if n < 0:
        return int(n)
    return dict(qry_aln=aq.raw[:n].decode(), ref_aln=ar.raw[:n].decode(), score=score.value, hit_boundary=bool(hb.value),
                band_width=bw.value, attempts=att.value)


def map_variations(ref, qry, mean_shift, band_width, extra_band_width=5, max_alignment_attempts=4):
    """-> dict(subs=[(pos, chr)], dels=[(pos, len)], inss=[(pos, seq)], hit_boundary, attempts) or an int error code"""
    lib = kswref.load_oracle()
    r, q = ref.encode() if isinstance(ref, str) else ref, qry.encode() if isinstance(qry, str) else qry
    rl, ql = len(r), len(q)
    sub_pos, sub_chr = np.zeros(rl + 1, np.int32), C.create_string_buffer(rl + 1)
    del_pos, del_len = np.zeros(rl + 3, np.int32), np.zeros(rl + 3, np.int32)
    ins_pos, ins_off, ins_len = np.zeros(ql + 2, np.int32), np.zeros(ql + 2, np.int32), np.zeros(ql + 2, np.int32)
    ins_seq = C.create_string_buffer(ql + 2)
    ns, nd, ni, hb, att = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int(0), C.c_int(0)
    P = lambda a: C.c_void_p(a.ctypes.data)
    rc = lib.orc_map_variations(r, rl, q, ql, mean_shift, band_width, extra_band_width, max_alignment_attempts, C.byref(ns), P(sub_pos),
                                sub_chr, C.byref(nd), P(del_pos), P(del_len), C.byref(ni), P(ins_pos), P(ins_off), P(ins_len), ins_seq,
                                C.byref(hb), C.byref(att))
    if rc != 0:
        return int(rc)
    return dict(subs=[(int(sub_pos[i]), sub_chr.raw[i:i + 1].decode()) for i in range(ns.value)],
                dels=[(int(del_pos[i]), int(del_len[i])) for i in range(nd.value)],
                inss=[(int(ins_pos[i]), ins_seq.raw[int(ins_off[i]):int(ins_off[i]) + int(ins_len[i])].decode()) for i in range(ni.value)],
                hit_boundary=bool(hb.value), attempts=att.value)


def apply_edit(ref, e):
    """Edit::apply of the reference (packages/pangraph/src/pangraph/edits.rs): the query the edit describes."""
    seq = list(ref)
    for pos, c in e["subs"]:
        seq[pos] = c
    for pos, ln in e["dels"]:
        for i in range(pos, pos + ln):
            seq[i] = ""
    out = []
    ins_at = {}
    for pos, s in e["inss"]:
        ins_at[pos] = ins_at.get(pos, "") + s
    for i in range(len(ref) + 1):
        if i in ins_at:
            out.append(ins_at[i])
        if i < len(ref):
            out.append(seq[i])
    return "".join(out)
