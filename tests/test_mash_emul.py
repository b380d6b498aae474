"""K7's tile schedule and per-position decisions (pangraph_b200/csrc/mash_core.h) replayed on the CPU (tests/mash_emul.cpp,
test-only) against the sequential restatement of the reference's minimizers_sketch (oracle/guide_tree_oracle.c): the SET of
(value, position) pairs the scan appends -- the reference's list repeats elements and is not in position order, and its only
consumer reduces it to sets (see mash_core.h)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gtref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(ROOT, "tests", "_build", "libmash_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tests", "mash_emul.cpp")
    deps = [src, os.path.join(ROOT, "pangraph_b200", "csrc", "mash_core.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.mash_emul_sketch.restype = C.c_int64
    return lib


def emul_sketch(lib, seq, sid, k, w):
    s = seq.encode() if isinstance(seq, str) else bytes(seq)
    cap = len(s) + 1
    val, pos = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
    n = lib.mash_emul_sketch(s, C.c_int64(len(s)), C.c_uint64(sid), k, w, C.c_void_p(val.ctypes.data), C.c_void_p(pos.ctypes.data),
                             C.c_int64(cap))
    assert 0 <= n <= cap, n
    got = [(int(val[i]), int(pos[i])) for i in range(n)]
    assert [p for _, p in got] == sorted(set(p for _, p in got))  # position order, each position once
    return got


def rand_seq(rng, n, alphabet=b"ACGT", n_frac=0.0, repeat=None):
    s = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=n)].copy()
    if repeat:
        unit = s[:repeat].copy()
        for st in range(0, n - repeat, repeat * 3):
            s[st:st + repeat] = unit
    if n_frac > 0:
        s[rng.random(n) < n_frac] = ord("N")
        if n > 60:
            st = int(rng.integers(0, n - 50))
            s[st:st + int(rng.integers(1, 50))] = ord("N")
    return s.tobytes()


def check(lib, seq, k, w, sid=3):
    want = gtref.mash_sketch(seq, sid, k=k, w=w)
    got = emul_sketch(lib, seq, sid, k, w)
    assert set(got) == set(want), (len(seq), k, w, sorted(set(got) ^ set(want))[:4])


def test_reference_vector(emul):
    seq = "CGATCCTTCGGGAACGTGTGACGCGAAGGTGCATGGGAGATCTCGCATTGCTGTTCTGGACGACGCGAAGAGTACTGCTACTTTCATGTCGCCTACGCCT"
    want = [(9685, 4294967328), (7669, 4294967355), (5583, 4294967359), (3600, 4294967386), (2383, 4294967415),
            (4791, 4294967427), (5338, 4294967451), (2190, 4294967461), (378, 4294967466)]
    assert emul_sketch(emul, seq, 1, 8, 16) == want  # no repeated values here: the list itself


@pytest.mark.parametrize("k,w", [(15, 100), (8, 16), (3, 5), (2, 3), (1, 1), (4, 255), (31, 7), (16, 40), (5, 1), (10, 100)])
def test_random_and_degenerate_sequences(emul, k, w):
    rng = np.random.default_rng(1000 * k + w)
    seqs = [rand_seq(rng, 3000), rand_seq(rng, 2500, n_frac=0.01), rand_seq(rng, 2000).lower(), rand_seq(rng, 4000, repeat=37),
            rand_seq(rng, 900, n_frac=0.2), b"A" * 700, b"ACGT" * 200, b"AT" * 300, rand_seq(rng, 3000, alphabet=b"AC"),
            rand_seq(rng, 3000, alphabet=b"ACGTUacgtuNRYKM-"), b"N" * 100, b"G", b"",
            rand_seq(rng, max(1, w + k - 2)), rand_seq(rng, w + k - 1), rand_seq(rng, w + k), rand_seq(rng, k), rand_seq(rng, max(1, k - 1))]
    for s in seqs:
        check(emul, s, k, w)


@pytest.mark.parametrize("k,w", [(15, 100), (8, 16), (3, 50), (6, 255)])
def test_tile_boundaries(emul, k, w):
    """lengths around multiples of the 4096-position tile; ambiguous bases and repeats (equal values inside every window:
    the oldest-wins rule and its duplicates) around the boundaries"""
    rng = np.random.default_rng(7 * k + w)
    for n in (4095, 4096, 4097, 8191, 8192, 8193, 4096 + w, 4096 + w + k, 4096 - w, 4096 + w + k - 1, 3 * 4096 + 17):
        check(emul, rand_seq(rng, n), k, w)
    s = bytearray(rand_seq(rng, 5 * 4096 + 100))
    for b in (4096, 8192, 12288, 16384):
        for d in (-w - k, -w, -k, -1, 0, 1, k - 1, w, w + k - 1):
            if 0 <= b + d < len(s):
                s[b + d] = ord("N")
    check(emul, bytes(s), k, w)
    check(emul, rand_seq(rng, 31) * 300, k, w)
    check(emul, b"A" * 9000, k, w)
    check(emul, rand_seq(rng, 30000, n_frac=0.002), k, w)
    check(emul, rand_seq(rng, 20000, alphabet=b"AC"), k, w)


def test_distance_from_sets(emul):
    """mash_distance over the emulated sets equals the restatement's matrix (mash_distance.rs only looks at the sets)"""
    from pangraph_b200 import synth
    gs = [g for _, g in synth.genomes(6, length=30_000, n_rearr=2, len_lo=300, len_hi=3000)]
    want = gtref.mash_distance(gs)
    sets = [set(v for v, _ in emul_sketch(emul, g, i, 15, 100)) for i, g in enumerate(gs)]
    n = len(gs)
    got = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            got[i, j] = got[j, i] = 1.0 - len(sets[i] & sets[j]) / len(sets[i])
    assert np.array_equal(got, want)


def test_window_and_kmer_size_sweep(emul):
    """window sizes around 1, the warp width and the block boundaries of the window-minimum scans, k from 1 to 31, on random,
    two-letter (many equal values per window), one-letter and N-rich sequences, the last two longer than a tile"""
    rng = np.random.default_rng(5)
    for w in (1, 2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 254, 255):
        for k in (1, 2, 3, 8, 15, 16, 31):
            for alphabet, n in ((b"ACGT", 700), (b"AC", 900), (b"A", 300), (b"ACGTN", 800), (b"ACGT", 4096 + w + 3), (b"AC", 8192 + 5)):
                check(emul, rand_seq(rng, n, alphabet=alphabet), k, w, sid=2)
