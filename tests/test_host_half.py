"""The host half of the path (split_matches / alignment_energy2 / filter_matches): the reference's own unit-test
vectors, checked against BOTH the Python oracle (oracle/host_half.py) and the product's C++ (libpgmm_b200.so; these
entry points are host code and run without a GPU), plus oracle-vs-product agreement on random CIGARs."""
import random

import pytest

from oracle import host_half as hh

CG = "3I 6M 3I 3M 4D 5M 14I 7M 3D 4I 5M 5D 3M 3I"


def aln(qry, ref, cigar, reverse, matches=0, length=0, quality=10, divergence=0.1, align=None):
    return dict(qry=qry, ref=ref, matches=matches, length=length, quality=quality, reverse=reverse,
                cigar=hh.parse_cigar(cigar), divergence=divergence, align=align)


# packages/pangraph/src/pangraph/split_matches.rs:271-594 (threshold 10)
SPLIT_KATS = [
    ("simple_case_forward", aln((0, 500, 200, 255), (1, 500, 100, 140), CG, False),
     [aln((0, 500, 203, 220), (1, 500, 100, 118), "6M3I3M4D5M", False, 14, 21),
      aln((0, 500, 234, 253), (1, 500, 118, 141), "7M3D4I5M5D3M", False, 15, 27)]),
    ("simple_case_reverse", aln((0, 500, 200, 256), (1, 500, 100, 141), CG, True),
     [aln((0, 500, 236, 253), (1, 500, 100, 118), "6M3I3M4D5M", True, 14, 21),
      aln((0, 500, 203, 222), (1, 500, 118, 141), "7M3D4I5M5D3M", True, 15, 27)]),
    ("side_patches_forward", aln((0, 257, 200, 257), (1, 56, 0, 56), "3I 3D 6M 3I 3M 4D 5M 14I 7M 3D 4I 5M 5D 3M 4I 12D", False, 29, 84),
     [aln((0, 257, 203, 220), (1, 56, 0, 21), "3D6M3I3M4D5M", False, 14, 24),
      aln((0, 257, 234, 257), (1, 56, 21, 44), "7M3D4I5M5D3M4I", False, 15, 31)]),
    ("side_patches_reverse_qry_leading", aln((0, 257, 200, 257), (1, 49, 0, 49), "3I 3D 6M 3I 3M 4D 5M 14I 7M 3D 4I 5M 5D 3M 4I 5D", True, 29, 77),
     [aln((0, 257, 237, 257), (1, 49, 0, 21), "3I3D6M3I3M4D5M", True, 14, 27),
      aln((0, 257, 204, 223), (1, 49, 21, 49), "7M3D4I5M5D3M5D", True, 15, 32)]),
    ("side_patches_reverse_qry_trailing", aln((0, 257, 0, 57), (1, 49, 0, 49), "3I 3D 6M 3I 3M 4D 5M 14I 7M 3D 4I 5M 5D 3M 4I 5D", True, 29, 77),
     [aln((0, 257, 37, 54), (1, 49, 0, 21), "3D6M3I3M4D5M", True, 14, 24),
      aln((0, 257, 0, 23), (1, 49, 21, 49), "7M3D4I5M5D3M5D4I", True, 15, 36)]),
]


def norm(a):
    d = dict(a)
    d["qry"], d["ref"] = tuple(int(v) for v in a["qry"]), tuple(int(v) for v in a["ref"])
    d["cigar"] = [(int(n), op) for n, op in a["cigar"]]
    d["matches"], d["length"], d["quality"] = int(a["matches"]), int(a["length"]), int(a["quality"])
    d["align"] = None
    return d


def test_keep_groups_reference_vector():
    """split_matches.rs:253-269"""
    cig = hh.parse_cigar("10I 20D 10M 20I 190D   40M 1D 1I 40M 1I 40M   1D 100I   200M 60I 60D 140M   200D   40M 2I 70M")
    assert hh.keep_groups(cig, 100) == [(5, 10), (13, 16), (18, 20)]


@pytest.mark.parametrize("name,inp,expected", SPLIT_KATS, ids=[k[0] for k in SPLIT_KATS])
def test_split_matches_reference_vectors(name, inp, expected):
    from pangraph_b200 import abi
    assert [norm(a) for a in hh.split_matches(inp, 10)] == [norm(e) for e in expected]
    got = abi.split_matches(inp, abi.alignment_args(indel_len_threshold=10))
    assert [norm(a) for a in got] == [norm(e) for e in expected]


def test_split_matches_rejects_clips():
    from pangraph_b200 import abi
    bad = aln((0, 500, 0, 300), (1, 500, 0, 300), "100M5S200M", False)
    with pytest.raises(ValueError):
        hh.split_matches(bad, 10)
    with pytest.raises(ValueError):
        abi.split_matches(bad, abi.alignment_args(indel_len_threshold=10))


def test_alignment_energy2_reference_vector():
    """energy.rs:90-112: -12.0"""
    from pangraph_b200 import abi
    a = aln((3, 100, 0, 50), (4, 200, 120, 200), "10I40M10D", False, 40, 60, 100, 0.02, 0.1)
    assert hh.alignment_energy2(a, 10.0, 10.0) == -12.0
    assert abi.alignment_energy2(a, abi.alignment_args(alpha=10.0, beta=10.0)) == -12.0
    no_div = dict(a, divergence=None)
    assert abi.alignment_energy2(no_div, abi.alignment_args(alpha=10.0, beta=10.0)) == hh.alignment_energy2(no_div, 10.0, 10.0) == -20.0


def test_filter_matches_reference_vector():
    """graph_merging.rs:301-375: the lower-divergence match wins the shared block; order by energy"""
    from pangraph_b200 import abi
    a0 = aln((0, 500, 100, 200), (1, 500, 200, 300), "100M", False, 100, 0, 0, 0.05)
    a1 = aln((2, 500, 100, 200), (3, 500, 200, 300), "100M", False, 100, 0, 0, 0.02)
    a2 = aln((2, 500, 150, 250), (4, 500, 200, 300), "100M", False, 100, 0, 0, 0.05)
    a3 = aln((5, 500, 100, 200), (6, 500, 200, 300), "100M", False, 100, 0, 0, 0.1)
    want = [norm(a1), norm(a0)]
    assert [norm(a) for a in hh.filter_matches([a0, a1, a2, a3], 10.0, 10.0)] == want
    assert [norm(a) for a in abi.filter_matches([a0, a1, a2, a3], abi.alignment_args(alpha=10.0, beta=10.0))] == want


def test_is_match_compatible_reference_vectors():
    """graph_merging.rs:254-299, through filter_matches: intervals that touch do not overlap, intervals that nest do"""
    from pangraph_b200 import abi
    args = abi.alignment_args(alpha=0.0, beta=0.0)
    first = [aln((0, 1000, 100, 200), (1, 1000, 200, 300), "100M", True, 100, 100), aln((0, 1000, 300, 400), (1, 1000, 400, 500), "100M", True, 100, 100)]
    ok = aln((0, 1000, 210, 290), (1, 1000, 310, 390), "80M", True, 80, 80, divergence=0.05)
    clash = aln((0, 1000, 310, 390), (1, 1000, 310, 390), "80M", True, 80, 80, divergence=0.05)
    assert len(abi.filter_matches(first + [ok], args)) == len(hh.filter_matches(first + [ok], 0.0, 0.0)) == 3
    assert len(abi.filter_matches(first + [clash], args)) == len(hh.filter_matches(first + [clash], 0.0, 0.0)) == 2


def random_alignment(rng, threshold):
    ops, prev = [], None
    for _ in range(rng.randint(1, 14)):
        op = rng.choice("MMMID=X")
        if op == prev:
            continue
        prev = op
        ops.append((rng.choice([1, 2, 5, threshold - 1, threshold, threshold + 3, 40]), op))
    ql = sum(n for n, op in ops if op in "MI=X")
    rl = sum(n for n, op in ops if op in "MD=X")
    qs, rs = rng.choice([0, 3, threshold, 50]), rng.choice([0, 2, threshold - 1, 70])
    qlen, rlen = qs + ql + rng.choice([0, 1, threshold - 1, 200]), rs + rl + rng.choice([0, 4, threshold, 90])
    return dict(qry=(rng.randint(0, 5), qlen, qs, qs + ql), ref=(rng.randint(6, 9), rlen, rs, rs + rl), matches=0, length=0,
                quality=rng.randint(0, 60), reverse=rng.random() < 0.5, cigar=ops,
                divergence=rng.choice([None, 0.0, 0.013, 0.2]), align=None)


def test_product_matches_oracle_on_random_cigars():
    from pangraph_b200 import abi
    rng = random.Random(11)
    for thr in (10, 100):
        args = abi.alignment_args(indel_len_threshold=thr, alpha=7.5, beta=3.25)
        pool = []
        for _ in range(300):
            a = random_alignment(rng, thr)
            want = hh.split_matches(a, thr)
            got = abi.split_matches(a, args)
            assert [norm(x) for x in got] == [norm(x) for x in want], a
            for x in want:
                assert abi.alignment_energy2(x, args) == hh.alignment_energy2(x, 7.5, 3.25)
            pool.extend(want)
        assert [norm(x) for x in abi.filter_matches(pool, args)] == [norm(x) for x in hh.filter_matches(pool, 7.5, 3.25)]
    assert abi.filter_matches([], abi.alignment_args()) == []
