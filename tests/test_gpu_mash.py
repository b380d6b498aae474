"""K7 (mash distance on the GPU) through the C-ABI against the restatement of the reference's mash_distance
(oracle/guide_tree_oracle.c): bit-identical float64 matrices -- all counts are integers, one IEEE division and subtraction
per entry -- on the reference's own vectors (PG/distance/mash/mash_distance.rs:91-151), genome families, degenerate inputs."""
import numpy as np
import pytest

import gtref
from test_oracle_guide_tree import MASH_FAMILY
from test_mash_emul import rand_seq

pytestmark = pytest.mark.gpu


def same(seqs, k=15, w=100):
    from pangraph_b200 import guide_tree as gt
    want = gtref.mash_distance(seqs, k=k, w=w)
    assert not isinstance(want, int), want
    got = gt.mash_distance(seqs, k=k, w=w)
    assert got.dtype == np.float64 and np.array_equal(got, want), (k, w, np.argwhere(got != want)[:4])
    return got


def test_reference_vectors():
    got = same(MASH_FAMILY, k=8, w=16)
    assert got[0, 1] == 1. - 6. / 9. and got[3, 4] == 0.625 and got[0, 3] == 1.0
    same([MASH_FAMILY[0], MASH_FAMILY[0]])
    same(["CATAGAAGCAGTCCCTGAGCACGACGCGTGTAACAATCGTTTTCAGACCTA"])


def test_errors_where_the_reference_panics():
    from pangraph_b200 import guide_tree as gt
    for seqs, kw in (([], {}), ([MASH_FAMILY[0], "ACGT"], {}), ([MASH_FAMILY[0], ""], {}), (["N" * 300, MASH_FAMILY[0]], {}),
                     ([MASH_FAMILY[0]], dict(k=32)), ([MASH_FAMILY[0]], dict(w=256)), ([MASH_FAMILY[0]], dict(k=0))):
        with pytest.raises(gt.GuideTreeError):
            gt.mash_distance(seqs, **kw)
    with pytest.raises(gt.GuideTreeError) as e:
        gt.mash_distance([MASH_FAMILY[0], MASH_FAMILY[1], "ACGT"])
    assert "sequence 2" in str(e.value)
    with pytest.raises(gt.GuideTreeError):  # k = 31 leaves 2 bits for the sequence number
        gt.mash_distance([MASH_FAMILY[0]] * 5, k=31, w=7)


@pytest.mark.parametrize("k,w", [(15, 100), (8, 16), (3, 5), (16, 40), (4, 255), (31, 7), (10, 100)])
def test_families_and_degenerate_sequences(k, w):
    from pangraph_b200 import synth
    rng = np.random.default_rng(100 * k + w)
    gs = [g for _, g in synth.genomes(7 if k < 31 else 4, length=60_000, n_rearr=3, len_lo=300, len_hi=5000)]
    same(gs, k, w)
    odd = [rand_seq(rng, 9000, n_frac=0.01), rand_seq(rng, 5000).lower(), rand_seq(rng, 8200, repeat=37), b"A" * 4097,
           b"ACGT" * 1100, rand_seq(rng, 4096, alphabet=b"AC"), rand_seq(rng, 4096 + w, alphabet=b"ACGTUacgtuNRYKM-"),
           rand_seq(rng, w + k), rand_seq(rng, 3 * 4096 + 17), rand_seq(rng, 31) * 300]
    if k == 31:
        odd = odd[:4]
    if not isinstance(gtref.mash_distance(odd, k=k, w=w), int):
        same(odd, k, w)
    same(gs[:2] + odd[:2], k, w)


def test_many_sequences_and_tree():
    """67 sequences (a pair block grid that is not a multiple of 4, several words of bitmap), then the tree: the joins the
    oracle takes on the oracle's matrix"""
    from pangraph_b200 import guide_tree as gt, synth
    gs = [g for _, g in synth.genomes(67, length=30_000, n_rearr=2, len_lo=300, len_hi=3000)]
    D, st = gt.mash_distance(gs, with_stats=True)
    want = gtref.mash_distance(gs)
    assert np.array_equal(D, want)
    assert st["bases"] == sum(len(g) for g in gs) and st["shared_values"] > 64 and st["minimizers"] >= st["unique_keys"] > 0
    t = gt.build_tree_using_neighbor_joining(gs, [f"g{i}" for i in range(67)])
    assert (t.left, t.right) == gtref.nj_tree(want)
    print("mash stats", st)


def test_genome_scale():
    """8 x 1 Mbp: the production parameters at a size where a tile-boundary or offset error cannot hide"""
    from pangraph_b200 import guide_tree as gt, synth
    gs = [g for _, g in synth.genomes(8, length=1_000_000)]
    D, st = gt.mash_distance(gs, with_stats=True)
    assert np.array_equal(D, gtref.mash_distance(gs))
    print("mash stats", st)
