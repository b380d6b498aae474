"""The nextalign / map_variations restatement (oracle/nextalign_oracle.c) against every unit-test vector the reference holds
for this path: packages/pangraph/src/align/map_variations.rs:185-366 (four Edits), align/nextclade/align_with_nextclade.rs:87-311
(four alignments incl. Ns, IUPAC-free edge case and the unalignable pair) and align/nextclade/align/align.rs:193-251 (band hit,
unaligned)."""
import naref

EXTRA = 5  # PangraphBuildArgs::default().extra_band_width (commands/build/build_args.rs:76)


def test_map_variations_vectors():
    cases = [
        ("ACTTTGCGTCTGATAGCTTAGCGGATATTTACTGTA", "ACTAGATTGAGTCTGATAGCTTAGCGGATATTGTA", -2, 3,
         dict(subs=[(6, "A")], dels=[(29, 4)], inss=[(3, "AGA")])),
        ("ACACTGATTTCGTCCCTTAGGTACTCTACACTGTAGCCTA", "CTGATTTAGTCCCTTAGGGGTTACTCTACACTGTAG", 2, 2,
         dict(subs=[(10, "A")], dels=[(0, 3), (36, 4)], inss=[(21, "GGT")])),
        ("ACACTGATTTCGTCCCTTAGGTACTCTACACTGTAGCCTA", "CCTGACACTGATTTAGTCCTAGGGGTTACTCTACACCGTAGCCTAGCCGCCG", -4, 2,
         dict(subs=[(10, "A"), (31, "C")], dels=[(15, 2)], inss=[(0, "CCTG"), (21, "GGT"), (40, "GCCGCCG")])),
        ("CGCCCTACTACAAGAGGGAACTTTTTTTTTAAGTATAGCCACAATAGCTGG", "CGCCCTACTACAAGAGGGAACGGGGGGGGGGGGGAAGTATAGCCACAATAGCTGG", -2, 11,
         dict(subs=[], dels=[(21, 9)], inss=[(21, "GGGGGGGGGGGGG")])),
    ]
    for r, q, ms, bw, want in cases:
        got = naref.map_variations(r, q, ms, bw, EXTRA, 4)
        assert {k: got[k] for k in ("subs", "dels", "inss")} == want, (r, got)
        assert naref.apply_edit(r, got) == q


def _aln(ref, qry, ms, bw, min_length=3, attempts=3):
    return naref.align_nuc_simplestripe(qry, ref, ms, bw, naref.params(min_length=min_length, max_alignment_attempts=attempts))


def test_align_with_nextclade_vectors():
    ref = "CTTGGAGGTTCCGTGGCTAGATAACAGAACATTCTTGGAATGCTGATCTTTATAAGCTCATGCGACACTTCGCATGGTGAGCCTTTGT"
    qry = "CTTGGAGGTTCCGTGGCTATAAAGATAACAGAACATTCTTGGAATGCTGATCAAGCTCATGGGACANNTCGCATGGTGGACAGCCTTTGT"
    a = _aln(ref, qry, 0, 4 + EXTRA)
    assert a["ref_aln"] == "CTTGGAGGTTCCGTGGCTA----GATAACAGAACATTCTTGGAATGCTGATCTTTATAAGCTCATGCGACACTTCGCATGGTG---AGCCTTTGT"
    assert not a["hit_boundary"]
    e = naref.map_variations(ref, qry, 0, 4, EXTRA, 3)
    assert e["subs"] == [(62, "G"), (67, "N"), (68, "N")] and e["dels"] == [(48, 5)] and e["inss"] == [(19, "TAAA"), (79, "GAC")]

    ref = "TGGTGCTGCAGCTTATTATGTGGNNNNNTTTTCTATTAAAATATAATGAAA"
    qry = "TGGTGCTGCAGCTTATTATGTGGAGGACTTTTCTATTAAAATATAATGAAA"
    a = _aln(ref, qry, 0, EXTRA)
    assert a["qry_aln"] == qry and a["ref_aln"] == ref and not a["hit_boundary"]
    e = naref.map_variations(ref, qry, 0, 0, EXTRA, 3)
    assert e["subs"] == [(23, "A"), (24, "G"), (25, "G"), (26, "A"), (27, "C")] and e["dels"] == [] and e["inss"] == []

    ref = "TGGTGCTGCNNNNNATTATGTGGGTTATCTTCAACCTTTTTTTAAAATATAATGAAAATGGAACCATTACAGATGCTNNNNNNNNTGCACTTGACCCTCTC"
    qry = "TGGTGCTGCAGCTTATTATGTGGGTTATCTTCAACCTTTTTTTAAAATATAATGAAAATGGAACCATTACAGATGCTGTAGACTGTGCACTTGACCCTCTC"
    a = _aln(ref, qry, 0, EXTRA)
    assert a["qry_aln"] == qry and a["ref_aln"] == ref
    e = naref.map_variations(ref, qry, 0, 0, EXTRA, 3)
    assert [p for p, _ in e["subs"]] == [9, 10, 11, 12, 13, 77, 78, 79, 80, 81, 82, 83, 84] and "".join(c for _, c in e["subs"]) == "AGCTTGTAGACTG"

    ref, qry = "A" * 37, "G" * 18  # unalignable: the whole reference deleted, the whole query inserted after the last base
    a = _aln(ref, qry, 70, EXTRA)
    assert a["ref_aln"] == "A" * 37 + "-" * 18 and a["qry_aln"] == "-" * 37 + "G" * 18 and not a["hit_boundary"]
    e = naref.map_variations(ref, qry, 70, 0, EXTRA, 3)
    assert e["subs"] == [] and e["dels"] == [(0, 37)] and e["inss"] == [(37, "G" * 18)]


def test_align_nuc_simplestripe_vectors():
    ref = "TTGGCCCCGGTGCTGTCCGTCAACACGTCGTCGTCCGGCGACCTACCTGGTCTCAAAGGAGGTTTTGTTAAATGAATTAGATGGGTAAGGTTACCACGTCA" + "A" * 31
    qry = "G" * 30 + "TTGGCCCCGGTGCTGTCCGTCAACACGTCGTCGTCCGGCGACCTACCTGGTCTCAAAGGAGGTTTTGTTAAATGAATTAGATGGGTAAGGTTACCACGTCA"
    p = naref.params(min_length=100, max_alignment_attempts=1)
    assert not naref.align_nuc_simplestripe(qry, ref, -30, 1, p)["hit_boundary"]
    assert not naref.align_nuc_simplestripe(qry, ref, 0, 31, p)["hit_boundary"]
    assert naref.align_nuc_simplestripe(qry, ref, 0, 30, p)["hit_boundary"]
    a = naref.align_nuc_simplestripe("G" * 18, "A" * 37, 70, 0, naref.params(min_length=3, max_alignment_attempts=1))
    assert a == dict(qry_aln="-" * 37 + "G" * 18, ref_aln="A" * 37 + "-" * 18, score=0, hit_boundary=False, band_width=0, attempts=1)


def test_errors_and_retry():
    assert naref.map_variations("ACGT", "", 0, 0) == -1          # query shorter than min_length 1
    assert naref.map_variations("ACGT", "ACxT", 0, 0) == -1      # to_nuc_seq rejects the character
    # a 40-base insertion with a band of 5: the band is doubled until the boundary is no longer hit
    import numpy as np
    rng = np.random.default_rng(3)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 400))
    qry = ref[:200] + "".join("ACGT"[i] for i in rng.integers(0, 4, 40)) + ref[200:]
    e = naref.map_variations(ref, qry, 0, 0, 5, 4)
    assert e["attempts"] > 1 and naref.apply_edit(ref, e) == qry
