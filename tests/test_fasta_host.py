"""FASTA input through the C-ABI (no GPU) against the reference's own unit tests (packages/pangraph/src/io/fasta.rs:300-934):
same records, same error messages; gzip and multi-file input (from_paths) on top; the bundled real data as read by the library
against a plain Python reading of the same files."""
import gzip
import os

import pytest

from pangraph_b200 import fasta
from pangraph_b200.fasta import FastaRecord as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def names_seqs(recs):
    return [(r.seq_name, r.desc, r.seq.decode(), r.index) for r in recs]


@pytest.mark.parametrize("data,want", [
    (b"", []), (b"\n \n \n\n", []),
    (b">seq1\nATCG\n", [("seq1", None, "ATCG", 0)]),
    (b"\n>seq1\nATCG\n", [("seq1", None, "ATCG", 0)]),
    (b"\n\n\n>seq1\nATCG\n", [("seq1", None, "ATCG", 0)]),
    (b">seq1\nATCG", [("seq1", None, "ATCG", 0)]),
    (b">seq1\nATCG\n>seq2\nGCTA\n", [("seq1", None, "ATCG", 0), ("seq2", None, "GCTA", 1)]),
    (b"\n>seq1\n\nATCG\n\n\n>seq2\nGCTA\n\n", [("seq1", None, "ATCG", 0), ("seq2", None, "GCTA", 1)]),
    (b">seq1\nATCG\n\n", [("seq1", None, "ATCG", 0)]),
    (b"\n\n>a\nACGCTCGATC\n\n>b\nCCGCGC", [("a", None, "ACGCTCGATC", 0), ("b", None, "CCGCGC", 1)]),
    (b">a\nACGCTCGATC\n>b\nCCGCGC\n>c", [("a", None, "ACGCTCGATC", 0), ("b", None, "CCGCGC", 1), ("c", None, "", 2)]),
    (b">a\nACGCTCGATC\n>b\n>c\nCCGCGC", [("a", None, "ACGCTCGATC", 0), ("b", None, "", 1), ("c", None, "CCGCGC", 2)]),
    (b">Identifier Description\nACGT\n>Identifier Description with spaces\nACGT\n\n\n",
     [("Identifier", "Description", "ACGT", 0), ("Identifier", "Description with spaces", "ACGT", 1)]),
    (b">\nACGT\n", [("", None, "ACGT", 0)]), (b"> \nACGT\n", [("", None, "ACGT", 0)]),
    (b">seq1\nACGTYRWSKMDVHBN\n", [("seq1", None, "ACGTYRWSKMDVHBN", 0)]),
    (b">a\nACGT\n>b\nGCTA\n>c\nTGCA\n", [("a", None, "ACGT", 0), ("b", None, "GCTA", 1), ("c", None, "TGCA", 2)]),
    (b">seq1\r\nAC\r\nGT\r\n", [("seq1", None, "ACGT", 0)]),
    (b">\n", []),  # the first record is_empty(): read_many stops there (fasta.rs:118-124)
])
def test_reader_vectors(data, want):
    assert names_seqs(fasta.read_many_str(data)) == want


def test_dedent_and_gap_alphabet():
    text = """>FluBuster-001
ACAGCCATGTATTG--
>CommonCold-AB
ACATCCCTGTA-TG--
>Ecoli/Joke/2024|XD
ACATCGCCNNA--GAC

>Sniffles-B
GCATCCCTGTA-NG--
>StrawberryYogurtCulture|\U0001F353
CCGGCCATGTATTG--
> SneezeC-19
CCGGCGATGTRTTG--
  >MisindentedVirus|D-skew
  TCGGCCGTGTRTTG--
"""
    want = [("FluBuster-001", None, "ACAGCCATGTATTG--", 0), ("CommonCold-AB", None, "ACATCCCTGTA-TG--", 1),
            ("Ecoli/Joke/2024|XD", None, "ACATCGCCNNA--GAC", 2), ("Sniffles-B", None, "GCATCCCTGTA-NG--", 3),
            ("StrawberryYogurtCulture|\U0001F353", None, "CCGGCCATGTATTG--", 4), ("", "SneezeC-19", "CCGGCGATGTRTTG--", 5),
            ("MisindentedVirus|D-skew", None, "TCGGCCGTGTRTTG--", 6)]
    assert names_seqs(fasta.read_many_str(text, fasta.DNA_WITH_GAP)) == want


def test_case_multiline_indentation():
    text = """>MixedCaseSeq
aCaGcCAtGtAtTG--
>LowercaseSeq
acagccatgtattg--
>UppercaseSeq
ACAGCCATGTATTG--
>MultilineSeq
ACAGCC
ATGT
ATTG--
>SkewedIndentSeq
  ACAGCC
ATGTATTG
 ATTG--
"""
    got = names_seqs(fasta.read_many_str(text, fasta.DNA_WITH_GAP))
    assert [g[2] for g in got] == ["ACAGCCATGTATTG--"] * 4 + ["ACAGCCATGTATTGATTG--"]
    assert names_seqs(fasta.read_many_str(b">seq1\nXYZ-xyz\n", "XYZ-")) == [("seq1", None, "XYZ-XYZ", 0)]
    assert names_seqs(fasta.read_many_str(b">seq1\nACGT-ACGT\n", fasta.DNA_WITH_GAP)) == [("seq1", None, "ACGT-ACGT", 0)]


@pytest.mark.parametrize("data,want", [
    (b"This is not a valid FASTA string.\nIt is not empty, and not entirely whitespace\nbut does not contain 'greater than' character.\n",
     "FASTA input is incorrectly formatted: expected at least one FASTA record starting with character '>', but none found"),
    (b">seq1\nACGT%ACGT\n", 'When processing sequence #1: ">seq1": FASTA input is incorrect: character "%" is not in the alphabet'),
    (b">seq1\nACGT-ACGT\n", 'When processing sequence #1: ">seq1": FASTA input is incorrect: character "-" is not in the alphabet'),
    (b">seq1\n%ACGT\n", 'When processing sequence #1: ">seq1": FASTA input is incorrect: character "%" is not in the alphabet'),
    (b">seq1\nACGT%\n", 'When processing sequence #1: ">seq1": FASTA input is incorrect: character "%" is not in the alphabet'),
    (b">a\nACGT\n>b some words\nAC GT\n", 'When processing sequence #2: ">b some words": FASTA input is incorrect: character " " is not in the alphabet'),
    (">a\nACéGT\n".encode(), 'When processing sequence #1: ">a": FASTA input is incorrect: character "é" is not in the alphabet'),
])
def test_reader_errors(data, want):
    with pytest.raises(fasta.FastaError) as e:
        fasta.read_many_str(data)
    assert str(e.value) == want


def test_files_gzip_and_concatenation(tmp_path):
    a, b, c = tmp_path / "a.fa", tmp_path / "b.fa.gz", tmp_path / "c.FA.GZ"
    a.write_bytes(b">a desc\nACGT\nacgt")          # no trailing newline: from_paths puts one between the files
    with gzip.open(b, "wb") as f:
        f.write(b">b\nGG\n")
    with open(c, "wb") as f:                         # two gzip members in one file (MultiGzDecoder reads both)
        f.write(gzip.compress(b">c1\nTT\n") + gzip.compress(b">c2\nCC\n"))
    got = names_seqs(fasta.read_many([str(a), str(b), str(c)]))
    assert got == [("a", "desc", "ACGTACGT", 0), ("b", None, "GG", 1), ("c1", None, "TT", 2), ("c2", None, "CC", 3)]
    assert names_seqs(fasta.read_many(str(b))) == [("b", None, "GG", 0)]
    with pytest.raises(fasta.FastaError) as e:
        fasta.read_many(str(tmp_path / "missing.fa"))
    assert "When opening file" in str(e.value)
    (tmp_path / "x.fa.zst").write_bytes(b"")
    with pytest.raises(fasta.FastaError) as e:
        fasta.read_many(str(tmp_path / "x.fa.zst"))
    assert "not supported" in str(e.value)


def test_files_on_disk():
    """the KAT pair of the alignment path as a plain file; and, where the reference's data directory is mounted (this
    container, not the GPU box), its 50-Mbp E. coli set through the gzip path against the genomes tests/golden holds"""
    path = os.path.join(ROOT, "tests", "golden", "kat_pair.fa")
    recs = fasta.read_many(path)
    want, name = [], None
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            name = line[1:].split(" ", 1)[0]
            want.append([name, ""])
        elif want:
            want[-1][1] += line.upper()
    assert [[r.seq_name, r.seq.decode()] for r in recs] == want and len(recs) == 2
    big = "/root/reference/data/ecoli.fa.gz"
    if os.path.exists(big):
        import realdata
        seqs, names = realdata.load_pair("ecoli")
        recs = fasta.read_many(big)
        assert len(recs) == 10 and sum(len(r.seq) for r in recs) == 49_434_912
        assert [r.seq_name for r in recs[:2]] == names and [r.seq for r in recs[:2]] == seqs
        assert [r.index for r in recs] == list(range(10))
