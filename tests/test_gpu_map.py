"""The whole path through the C-ABI (mm_idx_str / mm_mapopt_update / mm_map / pgmm_map_batch) against the reference:
every field of every mm_reg1_t, every CIGAR, in the reference's order."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_fa(path):
    recs = []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            recs.append([line[1:], ""])
        elif line:
            recs[-1][1] += line
    return recs


def compare(seqs, names, preset="asm10", k=None, per_query=False):
    from oracle import refmm2
    from pangraph_b200 import abi
    want, mid_ref = refmm2.ref_map_all(seqs, names, preset, k, 90, threads=8)
    idx = abi.Index(seqs, names, preset, k, 90)
    assert idx.mo.mid_occ == mid_ref
    got = idx.map_batch()
    if per_query:
        single = [idx.map_one(s, n) for s, n in zip(seqs, names)]
        assert single == got
    idx.close()
    assert [len(g) for g in got] == [len(w) for w in want]
    for qi, (g, w) in enumerate(zip(got, want)):
        for ri, (a, b) in enumerate(zip(g, w)):
            assert a == b, (qi, ri, a[:18], b[:18])
    return sum(len(w) for w in want)


def test_reference_golden_vector():
    """The reference's own boundary KAT (align_with_minimap2_lib.rs:135-204): asm20, k=10."""
    from pangraph_b200 import abi
    recs = read_fa(os.path.join(ROOT, "tests", "golden", "kat_pair.fa"))
    idx = abi.Index([s for _, s in recs], [n for n, _ in recs], "asm20", 10, 90)
    got = idx.map_batch()
    idx.close()
    assert got[0] == []  # query "1" > target "0" in strcmp order: skipped by MM_F_NO_DUAL
    (r,) = got[1]
    assert (r[2], r[4], r[5], r[6], r[7], r[11], r[12]) == (0, 0, 996, 0, 998, 969, 998)
    assert (r[15] >> 10) & 1 == 0 and r[15] & 0xff == 0  # forward strand, mapq 0
    cap, dp_score, dp_max, dp_max2, n_ambi, cig, de = r[18]
    assert dp_score == 845 and de == 0.029058116232464903
    assert "".join(f"{c >> 4}{'MIDNSHP=XB'[c & 15]}" for c in cig) == "545M1D225M1D226M"
    n = compare([s for _, s in recs], [n for n, _ in recs], "asm20", 10, per_query=True)
    assert n == 1


@pytest.mark.parametrize("preset", ["asm5", "asm10", "asm20"])
def test_small_genomes_all_vs_all(preset):
    """6 x 60 kbp related genomes with rearrangements, decimal block-id names as pangraph passes them."""
    from pangraph_b200 import synth
    gs = synth.genomes(6, length=60_000, n_rearr=6, len_lo=300, len_hi=8000)
    names = [str(v) for v in (3, 17, 5, 10442385907364519937, 100, 9)]
    n = compare([g for _, g in gs], names, preset, per_query=(preset == "asm10"))
    assert n > 5


def test_repeats_self_hits_and_ambiguous_bases():
    """Internal repeats (self hits, anchor ties, high-occurrence seeds), Ns, lower case, a reverse-complemented genome."""
    from pangraph_b200 import synth
    rng = np.random.default_rng(5)
    anc = synth.ancestor(80_000, 7)
    unit = anc[1000:3500].copy()
    for st in (9000, 20000, 41000, 66000):
        anc[st:st + len(unit)] = unit
    short = anc[500:560].copy()
    for st in range(30000, 36000, 60):  # tandem array: minimizers above mid_occ
        anc[st:st + 60] = short
    gs = [synth.mutate(anc, 900 + i, n_rearr=4, len_lo=300, len_hi=6000) for i in range(4)]
    gs[1][5000:5040] = ord("N")
    gs[2] = synth.revcomp(gs[2])
    seqs = [g.tobytes() for g in gs]
    seqs[3] = seqs[3].lower()
    n = compare(seqs, ["0", "1", "2", "3"], "asm10", per_query=True)
    assert n > 10


def test_one_megabase_pair():
    from pangraph_b200 import synth
    gs = synth.genomes(2, length=1_000_000)
    n = compare([g for _, g in gs], ["0", "1"], "asm10")
    assert n >= 1


def test_many_short_blocks():
    """Upper merge levels: hundreds of short blocks, most of them unrelated."""
    from pangraph_b200 import synth
    rng = np.random.default_rng(3)
    anc = synth.ancestor(40_000, 11)
    seqs, names = [], []
    for i in range(120):
        a, b = sorted(int(v) for v in rng.integers(0, 40_000, size=2))
        if b - a < 150:
            b = a + 150
        seg = synth.mutate(anc[a:b].copy(), 50 + i, n_rearr=0)
        if i % 3 == 0:
            seg = synth.revcomp(seg)
        seqs.append(seg.tobytes())
        names.append(str(int(rng.integers(0, 2**63))))
    seqs.append(b"ACGT")  # shorter than a k-mer
    names.append("7")
    n = compare(seqs, names, "asm10")
    assert n > 50


def test_rounds_without_the_dp_service_and_with_the_chain_service(ref):
    """PGMM_DP_SERVICE=0 keeps every round on its own DP engine (one arena and one set of streams per context) and
    PGMM_CHAIN_SERVICE=1 sends the chaining fill through the cross-round chain service: the same hits as with the
    defaults (DP service on, chain fill on the round's own stream), which every other test in this file uses."""
    import subprocess
    import sys
    code = (
        "import os, sys; sys.path.insert(0, '.');"
        "from pangraph_b200 import abi, synth; from oracle import refmm2;"
        "gs = synth.genomes(3, length=60_000, n_rearr=5, len_lo=300, len_hi=6000);"
        "seqs, names = [g for _, g in gs], ['0', '1', '2'];"
        "idx = abi.Index(seqs, names, 'asm10', None, 90); got = idx.map_batch(); idx.close();"
        "want, _ = refmm2.ref_map_all(seqs, names, 'asm10', None, 90);"
        "assert got == want and sum(len(g) for g in got) > 0; print('ok')"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PGMM_DP_SERVICE="0", PGMM_CHAIN_SERVICE="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
