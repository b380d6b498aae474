"""The band-coordinate formulation of K6 (pangraph_b200/csrc/nextalign_core.h + nextalign_host.cpp) on the CPU: a serial
emulation of the kernel's lane schedule (tests/na_emul.cpp, test-only) against the oracle restatement of the reference
(oracle/nextalign_oracle.c) -- the reference's unit vectors, degenerate bands (left / right of the matrix, width 0..), retries,
IUPAC codes and Ns, empty reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import naref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(ROOT, "tests", "_build", "libna_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "na_emul.cpp"), os.path.join(ROOT, "pangraph_b200", "csrc", "nextalign_host.cpp")]
    deps = srcs + [os.path.join(ROOT, "pangraph_b200", "csrc", f) for f in ("nextalign_core.h", "nextalign_host.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-o", out] + srcs, check=True)
    return C.CDLL(out)


def emul_map_variations(lib, ref, qry, ms, bw, extra=5, attempts=4):
    r, q = ref.encode(), qry.encode()
    rl, ql = len(r), len(q)
    sub_pos, sub_chr = np.zeros(rl + 1, np.int32), C.create_string_buffer(rl + 1)
    del_pos, del_len = np.zeros(rl + 3, np.int32), np.zeros(rl + 3, np.int32)
    ins_pos, ins_len, ins_seq = np.zeros(ql + 2, np.int32), np.zeros(ql + 2, np.int32), C.create_string_buffer(ql + 2)
    ns, nd, ni, hb, att, sc = (C.c_int32(0) for _ in range(6))
    P = lambda a: C.c_void_p(a.ctypes.data)
    rc = lib.na_emul_map_variations(r, rl, q, ql, ms, bw, extra, attempts, C.byref(ns), P(sub_pos), sub_chr, C.byref(nd), P(del_pos),
                                    P(del_len), C.byref(ni), P(ins_pos), P(ins_len), ins_seq, C.byref(hb), C.byref(att), C.byref(sc))
    if rc != 0:
        return int(rc)
    inss, off = [], 0
    for k in range(ni.value):
        inss.append((int(ins_pos[k]), ins_seq.raw[off:off + int(ins_len[k])].decode()))
        off += int(ins_len[k])
    return dict(subs=[(int(sub_pos[i]), sub_chr.raw[i:i + 1].decode()) for i in range(ns.value)],
                dels=[(int(del_pos[i]), int(del_len[i])) for i in range(nd.value)], inss=inss, hit_boundary=bool(hb.value),
                attempts=att.value)


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))


def mutate(rng, ref, sub=0.03, indel=0.01, max_indel=12, iupac=0.0):
    out, i = [], 0
    while i < len(ref):
        x = rng.random()
        if x < indel / 2:
            i += int(rng.integers(1, max_indel))
        elif x < indel:
            out.append(rand_seq(rng, int(rng.integers(1, max_indel))))
        else:
            c = ref[i]
            y = rng.random()
            if y < sub:
                c = "ACGT"[int(rng.integers(0, 4))]
            elif y < sub + iupac:
                c = "NRYKMSWBDHV"[int(rng.integers(0, 11))]
            out.append(c)
            i += 1
    return "".join(out)


def check(lib, ref, qry, ms, bw, extra=5, attempts=4):
    want = naref.map_variations(ref, qry, ms, bw, extra, attempts)
    got = emul_map_variations(lib, ref, qry, ms, bw, extra, attempts)
    assert got == want, (len(ref), len(qry), ms, bw, extra, attempts, got if isinstance(got, int) else {k: got[k] for k in ("hit_boundary", "attempts")},
                         want if isinstance(want, int) else {k: want[k] for k in ("hit_boundary", "attempts")})
    return want


def test_reference_vectors(emul):
    for r, q, ms, bw in [("ACTTTGCGTCTGATAGCTTAGCGGATATTTACTGTA", "ACTAGATTGAGTCTGATAGCTTAGCGGATATTGTA", -2, 3),
                         ("ACACTGATTTCGTCCCTTAGGTACTCTACACTGTAGCCTA", "CTGATTTAGTCCCTTAGGGGTTACTCTACACTGTAG", 2, 2),
                         ("ACACTGATTTCGTCCCTTAGGTACTCTACACTGTAGCCTA", "CCTGACACTGATTTAGTCCTAGGGGTTACTCTACACCGTAGCCTAGCCGCCG", -4, 2),
                         ("CGCCCTACTACAAGAGGGAACTTTTTTTTTAAGTATAGCCACAATAGCTGG", "CGCCCTACTACAAGAGGGAACGGGGGGGGGGGGGAAGTATAGCCACAATAGCTGG", -2, 11),
                         ("A" * 37, "G" * 18, 70, 0), ("A" * 37, "G" * 18, -70, 0), ("ACGT", "", 0, 0), ("ACGT", "ACxT", 0, 0), ("", "ACGT", 0, 0)]:
        check(emul, r, q, ms, bw)


def test_random_pairs_and_degenerate_bands(emul):
    rng = np.random.default_rng(11)
    n_retry = 0
    for it in range(400):
        rl = int(rng.integers(1, 260))
        ref = rand_seq(rng, rl, "ACGT" if it % 7 else "ACGTN")
        kind = it % 5
        if kind == 0:
            qry = rand_seq(rng, int(rng.integers(1, 260)))  # unrelated
        else:
            qry = mutate(rng, ref, iupac=0.02 if kind == 2 else 0.0) or "A"
            if kind == 3:  # terminal overhangs
                qry = rand_seq(rng, int(rng.integers(0, 30))) + qry[int(rng.integers(0, 20)):]
            if kind == 4:
                qry = qry[:max(1, len(qry) - int(rng.integers(0, 40)))]
        ms = int(rng.integers(-40, 40)) if it % 3 == 0 else int(rng.integers(-3, 4)) if it % 3 == 1 else int(rng.integers(-300, 300))
        bw = int(rng.integers(0, 6)) if it % 2 else int(rng.integers(0, 80))
        extra = 5 if it % 4 else 0
        attempts = int(rng.integers(1, 5))
        w = check(emul, ref, qry, ms, bw, extra, attempts)
        if not isinstance(w, int):
            n_retry += w["attempts"] > 1
            assert naref.apply_edit(ref, w) == qry
    assert n_retry > 20


def test_wide_band_many_columns_per_lane(emul):
    rng = np.random.default_rng(12)
    ref = rand_seq(rng, 700)
    qry = ref[:300] + rand_seq(rng, 150) + ref[340:]
    for bw in (40, 100, 200):  # W = 91 .. 411: 3 .. 13 band columns per lane
        check(emul, ref, qry, -55, bw, 5, 2)
