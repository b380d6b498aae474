"""Committed golden vectors (tests/golden/golden_regs.json, generated from the unmodified reference by
tests/golden/make_golden.py): the reference's boundary KAT and a synthetic family with inversion rescues (5 hits) and
z-drop splits (48 hits).  CPU: the reference build, the oracle's host half and the product's host logic reproduce them.
GPU: the whole CUDA path reproduces them, hits and filtered matches."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_regs.json")))


def cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return {"kat": mg.kat_case(), "synth": mg.synth_case()}


def as_tuples(regs):
    return [[tuple(r[:18]) + (None if r[18] is None else (r[18][0], r[18][1], r[18][2], r[18][3], r[18][4], tuple(r[18][5]), r[18][6]),)
             for r in q] for q in regs]


def norm_matches(ms, cigar_to_str):
    out = []
    for m in ms:
        d = dict(m)
        d["qry"], d["ref"] = [int(v) for v in m["qry"]], [int(v) for v in m["ref"]]
        d["cigar"] = cigar_to_str(m["cigar"]) if not isinstance(m["cigar"], str) else m["cigar"]
        d["matches"], d["length"], d["quality"] = int(m["matches"]), int(m["length"]), int(m["quality"])
        out.append(d)
    return out


@pytest.mark.parametrize("name", ["kat", "synth"])
def test_reference_build_reproduces_golden(ref, name):
    from oracle import refmm2
    seqs, names = cases()[name]
    g = GOLD[name]
    regs, mid = refmm2.ref_map_all(seqs, names, g["preset"], g["k"], 90)
    assert mid == g["mid_occ"] and regs == as_tuples(g["regs"])


@pytest.mark.parametrize("name", ["kat", "synth"])
def test_oracle_host_half_reproduces_golden_matches(name):
    from oracle import host_half
    seqs, names = cases()[name]
    g = GOLD[name]
    ms = host_half.find_filtered_matches(as_tuples(g["regs"]), names, [len(s) for s in seqs])
    assert norm_matches(ms, host_half.cigar_str) == norm_matches(g["matches"], None)


@pytest.mark.parametrize("name", ["kat", "synth"])
def test_product_host_logic_reproduces_golden(ref, name):
    import hostlogic
    seqs, names = cases()[name]
    g = GOLD[name]
    got, mid = hostlogic.map_all(hostlogic.load(), seqs, names, g["preset"], g["k"], 90, threads=2)
    assert mid == g["mid_occ"] and got == as_tuples(g["regs"])


def test_kat_is_the_reference_unit_test():
    """align_with_minimap2_lib.rs:188-200: qry 0 [0,996) len 998, ref 1 [0,998) len 1000, +, 969/998, mapq 0, AS 845"""
    (m,) = GOLD["kat"]["matches"]
    assert m["qry"] == [0, 998, 0, 998] or m["qry"] == [0, 998, 0, 996]
    (r,) = [r for q in GOLD["kat"]["regs"] for r in q]
    assert (r[4], r[5], r[6], r[7], r[11], r[12], r[15] & 0xff) == (0, 996, 0, 998, 969, 998, 0)
    assert r[18][1] == 845 and r[18][6] == 0.029058116232464903
    assert "".join(f"{c >> 4}{'MIDNSHP=XB'[c & 15]}" for c in r[18][5]) == "545M1D225M1D226M"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat", "synth"])
def test_cuda_path_reproduces_golden(name):
    from oracle import host_half
    from pangraph_b200 import abi
    seqs, names = cases()[name]
    g = GOLD[name]
    idx = abi.Index(seqs, names, g["preset"], g["k"], 90)
    assert idx.mo.mid_occ == g["mid_occ"]
    assert idx.map_batch() == as_tuples(g["regs"])
    assert idx.map_self() == as_tuples(g["regs"])
    idx.close()
    sens = {"asm5": 5, "asm10": 10, "asm20": 20}[g["preset"]]
    args = abi.alignment_args(sensitivity=sens, kmer_length=g["k"])
    blocks = {int(n): s for n, s in zip(names, seqs)}
    got = abi.find_filtered_matches(blocks, args)
    assert norm_matches(got, host_half.cigar_str) == norm_matches(g["matches"], None)


def test_host_chainer_against_chain_fixture():
    """The product's host chainer (segments + AVL arbiter + backtrack) on the committed mg_lchain_rmq fixtures
    (tests/golden/golden_chain.npz, made by make_golden_chain.py from the unmodified reference)."""
    import ctypes as C

    import numpy as np

    import hostlogic
    hl = hostlogic.load()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_chain.npz"))
    p = g["params"]
    hl.pgmm_test_chain_rmq.restype = C.c_int64
    for name in ("noisy", "colinear", "repeat_array"):
        a = g[name + "_in"].copy()
        u = np.zeros(len(a) + 1, dtype=np.uint64)
        n_a = C.c_int64(0)
        n_u = hl.pgmm_test_chain_rmq(C.c_void_p(a.ctypes.data), C.c_int64(len(a)), int(p[0]), int(p[1]), int(p[2]), int(p[3]), int(p[4]),
                                     int(p[5]), int(p[6]), C.c_float(p[7]), C.c_float(p[8]), C.c_void_p(u.ctypes.data), C.byref(n_a))
        assert n_u == len(g[name + "_u"]) and n_a.value == len(g[name + "_kept"]), name
        assert np.array_equal(u[:n_u], g[name + "_u"]) and np.array_equal(a[:n_a.value], g[name + "_kept"]), name
