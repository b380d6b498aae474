"""The product's HOST logic (anchor order, RMQ chaining, hit bookkeeping, wave-scheduled DP stitching, filters, mapq)
on CPU: tests/_build/libpgmm_hostlogic.so = the product's host sources + a backend that forwards the device stages to
the reference's own C.  Every mm_reg1_t field and CIGAR must equal what the reference's mm_map returns."""
import gzip
import os

import numpy as np
import pytest

import hostlogic

REF_DATA = "/root/reference"


@pytest.fixture(scope="module")
def hl():
    return hostlogic.load()


def read_fa(path, limit=None):
    op = gzip.open if path.endswith(".gz") else open
    recs = []
    with op(path, "rt") as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if limit and len(recs) >= limit:
                    break
                recs.append([line[1:].split()[0], []])
            elif recs:
                recs[-1][1].append(line.upper())
    return [(n, "".join(s)) for n, s in recs]


def check(hl, seqs, names, preset="asm10", k=None, threads=4):
    from oracle import refmm2
    want, mid = refmm2.ref_map_all(seqs, names, preset, k, 90, threads=8)
    got, mid2 = hostlogic.map_all(hl, seqs, names, preset, k, 90, threads=threads)
    assert mid == mid2
    assert got == want
    return sum(len(w) for w in want)


def test_reference_golden_vector(ref, hl):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    recs = read_fa(os.path.join(root, "tests", "golden", "kat_pair.fa"))
    assert check(hl, [s for _, s in recs], [n for n, _ in recs], "asm20", 10) == 1


@pytest.mark.parametrize("preset", ["asm5", "asm10", "asm20"])
def test_synthetic_family(ref, hl, preset):
    from pangraph_b200 import synth
    gs = synth.genomes(5, length=50_000, n_rearr=6, len_lo=300, len_hi=8000)
    names = [str(v) for v in (3, 17, 5, 10442385907364519937, 100)]
    assert check(hl, [g for _, g in gs], names, preset) > 4


def test_repeats_inversions_and_splits(ref, hl):
    """Long inversions inside chains (z-drop split + inversion rescue), repeats (self hits, ties), Ns."""
    from pangraph_b200 import synth
    anc = synth.ancestor(120_000, 3)
    unit = anc[2000:5000].copy()
    for st in (30000, 61000, 99000):
        anc[st:st + len(unit)] = unit
    gs = [synth.mutate(anc, 70 + i, n_rearr=8, len_lo=500, len_hi=9000) for i in range(3)]
    gs[1][7000:7030] = ord("N")
    n = check(hl, [g.tobytes() for g in gs], ["0", "1", "2"], "asm10", threads=1)
    assert n > 10


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DATA, "data", "russian_doll_plasmids.fa.gz")), reason="reference data not mounted")
def test_reference_bundled_plasmids(ref, hl):
    recs = read_fa(os.path.join(REF_DATA, "data", "russian_doll_plasmids.fa.gz"))
    assert check(hl, [s for _, s in recs], [str(i) for i in range(len(recs))]) == 18
    recs = read_fa(os.path.join(REF_DATA, "packages", "pypangraph", "tests", "data", "plasmids.fa.gz"), limit=6)
    assert check(hl, [s for _, s in recs], [str(i * 7919) for i in range(len(recs))]) > 100


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DATA, "data", "ecoli.fa.gz")), reason="reference data not mounted")
def test_reference_bundled_ecoli_pair(ref, hl):
    """BASELINE config 2 at full size on the CPU: the first two genomes of data/ecoli.fa.gz (4.6 + 4.8 Mbp, 494 k anchors,
    ~36 k DP problems, inversion rescues, z-drop splits): all 2002 hits identical to the reference's."""
    recs = read_fa(os.path.join(REF_DATA, "data", "ecoli.fa.gz"), limit=2)
    assert check(hl, [s for _, s in recs], ["0", "1"], threads=8) == 2002


def test_wave_scheduler_variants(ref, hl, monkeypatch):
    """The two latency measures of the DP scheduler -- the exact second pass of a long fill queued with its first pass,
    hits finished in the wave after which nothing is pending for them -- change when work is done, not what comes out:
    the same hits with both on (default, every other test), both off, and with every fill speculated."""
    from pangraph_b200 import synth
    anc = synth.ancestor(150_000, 5)
    gs = [synth.mutate(anc, 900 + i, n_rearr=10, len_lo=500, len_hi=12000).tobytes() for i in range(3)]
    for early, spec in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")):
        monkeypatch.setenv("PGMM_EARLY_FINISH", early)
        monkeypatch.setenv("PGMM_SPEC_FILL_LEN", spec)
        assert check(hl, gs, ["0", "1", "2"], threads=2) > 10
