"""The guide-tree restatement (oracle/guide_tree_oracle.c) against the reference's own unit vectors
(PG = packages/pangraph/src): PG/distance/mash/hash.rs:20-27, minimizer.rs:190-219, mash_distance.rs:91-151,
PG/tree/neighbor_joining.rs:111-151 and the trees its two disabled tests expect (:203-288)."""
import numpy as np
import pytest

import gtref

INF = float("inf")

MASH_FAMILY = [
    "CATAGAAGCAGTCCCTGAGCACGACGCGTGTAACAATCGTTTTCAGACCTAGGACGTTAGAATATCGATCGCACGCTACGACCGACGATTAGCCGCACGAGCAAGTCGAAAACCCGAGTTAAGAGGCTGGACGTGATCCTAGACTTCGTC",
    "CATAGAAGCAGTCCCTGAGCACGAGGCGCGCAACAATCGTTTTCAGCCCTAGGACGTTAGAATATTGATCACAAGCTACGACCGACGATTAGCCGCACGAGCAAGTCGACAACCCGAGTTAAGAGGCTGGACGTGATGCTAGACTTCGTC",
    "CATAGAAGCAGTCCCTGAGCATGACGCGCGCAACGATCGTTTTCAGCCCTAGCACGTGAGAATATTGATCACAAGCTACGACCGACGATTAGCCGCACGAGCTAGTCGCCAACCCGAGTAAGGAGGCTGGACGTGATGCTAGACTACGTC",
    "ACATCAAAACTTAAAGTCGGTTACCATCTACAAATGTAGTAAGGGGGATTCTAATGAGAGAAGTGGACTGTGTAGATGGACCCGCTCACCTGCCCAGTATCTTAGTGGCGTATTCAGGATCTGGGAGGATTTGTTATTGCCTATTAGAGA",
    "ACATCAAAACTTAAAGTCGGTTCCCATCTACAAAAGTAGAAAGGGGGATTCTAATGAGAGATGTGGACTGTGTAGATGGACCCGCTAACCTGGCCAGTTTCTTAGTGGCTTAATCAGGATCTGGGAGGATTCGTTACTGCCTATTAGAGA",
    "ACATCAGAACTTAAAGTCGGTTCCTATCTCCAAAAGTATAAAGTGGGATTCTAATGAGAGATGTGGACTGTGTCGATAAACCCGCTAACCTGGCCTGTTTCTTGTTGGCTTAATCAGGATCTGAGAGGATTCGTTACTGCCTAGTAGTGA",
]

WIKI = np.array([[0.0, 5.0, 9.0, 9.0, 8.0],
                 [5.0, 0.0, 10.0, 10.0, 9.0],
                 [9.0, 10.0, 0.0, 8.0, 7.0],
                 [9.0, 10.0, 8.0, 0.0, 3.0],
                 [8.0, 9.0, 7.0, 3.0, 0.0]])


@pytest.mark.parametrize("x,mask,want", [(0, 0, 0), (123, 0, 0), (0, 456, 136), (123, 456, 384)])
def test_hash(x, mask, want):
    assert gtref.mash_hash(x, mask) == want


def test_minimizers_sketch_general_case():
    seq = "CGATCCTTCGGGAACGTGTGACGCGAAGGTGCATGGGAGATCTCGCATTGCTGTTCTGGACGACGCGAAGAGTACTGCTACTTTCATGTCGCCTACGCCT"
    want = [(9685, 4294967328), (7669, 4294967355), (5583, 4294967359), (3600, 4294967386), (2383, 4294967415),
            (4791, 4294967427), (5338, 4294967451), (2190, 4294967461), (378, 4294967466)]
    assert gtref.mash_sketch(seq, 1, k=8, w=16) == want


def test_minimizers_sketch_empty():
    assert gtref.mash_sketch("", 0) == []          # the reference's "No minimizers found for seq." error
    assert gtref.mash_sketch("ACGTACGTACGTAC", 0) == []  # shorter than k
    assert gtref.mash_sketch("N" * 500, 0) == []


def test_mash_distance_general_case():
    want = np.array([[0.0, 1. - 6. / 9., 0.75, 1.0, 1.0, 1.0],
                     [1. - 6. / 9., 0.0, 0.5, 1.0, 1.0, 1.0],
                     [0.75, 0.5, 0.0, 1.0, 1.0, 1.0],
                     [1.0, 1.0, 1.0, 0.0, 0.625, 0.875],
                     [1.0, 1.0, 1.0, 0.625, 0.0, 5. / 7.],
                     [1.0, 1.0, 1.0, 0.875, 5. / 7., 0.0]])
    got = gtref.mash_distance(MASH_FAMILY, k=8, w=16)
    assert np.array_equal(got, want)  # bit for bit, like the reference's assert_eq!


def test_mash_distance_edge_cases():
    assert gtref.mash_distance([]) == -1000000
    assert np.array_equal(gtref.mash_distance([MASH_FAMILY[0], MASH_FAMILY[0]]), np.zeros((2, 2)))
    assert np.array_equal(gtref.mash_distance(["CATAGAAGCAGTCCCTGAGCACGACGCGTGTAACAATCGTTTTCAGACCTA"]), np.zeros((1, 1)))
    assert gtref.mash_distance([MASH_FAMILY[0], "ACGT"]) == -2  # a sequence without minimizers: the reference panics


def test_create_q_matrix():
    want = np.array([[INF, -50.0, -38.0, -34.0, -34.0],
                     [-50.0, INF, -38.0, -34.0, -34.0],
                     [-38.0, -38.0, INF, -40.0, -40.0],
                     [-34.0, -34.0, -40.0, INF, -48.0],
                     [-34.0, -34.0, -40.0, -48.0, INF]])
    assert np.array_equal(gtref.nj_q_matrix(WIKI), want)


def test_dist():
    assert np.array_equal(gtref.nj_dist(WIKI, 0, 1), np.array([0., 0., 7., 7., 6.]))


def test_join_steps():
    """the D matrices neighbor_joining.rs:153-201 (test_join, disabled there) expects after one and two joins"""
    D = WIKI.copy()
    for want in ([[0.0, 7.0, 7.0, 6.0], [7.0, 0.0, 8.0, 7.0], [7.0, 8.0, 0.0, 3.0], [6.0, 7.0, 3.0, 0.0]],
                 [[0.0, 4.0, 3.0], [4.0, 0.0, 3.0], [3.0, 3.0, 0.0]]):
        Q = gtref.nj_q_matrix(D)
        r, c = np.unravel_index(np.argmin(Q), Q.shape)  # numpy's argmin: first minimum in row-major order as well
        i, j = min(r, c), max(r, c)
        dn = gtref.nj_dist(D, i, j)
        D[i, :], D[:, i], D[i, i] = dn, dn, 0.0
        D = np.delete(np.delete(D, j, 0), j, 1)
        assert np.array_equal(D, np.array(want))


def test_tree_wikipedia():
    left, right = gtref.nj_tree(WIKI)
    names = list("ABCDE")
    assert gtref.to_newick(left, right, names) == "((((A,B),C),D),E);"
    order = [names[v] if v < 5 else "" for v in gtref.postorder(left, right, 5)]
    assert order == ["A", "B", "", "C", "", "D", "", "E", ""]


def test_tree_eight_taxa():
    D = np.array([[0.0, 46.0, 37.0, 46.0, 46.0, 14.0, 37.0, 1.0],
                  [46.0, 0.0, 46.0, 7.0, 1.0, 46.0, 46.0, 46.0],
                  [37.0, 46.0, 0.0, 46.0, 46.0, 37.0, 1.0, 37.0],
                  [46.0, 7.0, 46.0, 0.0, 7.0, 46.0, 46.0, 46.0],
                  [46.0, 1.0, 46.0, 7.0, 0.0, 46.0, 46.0, 46.0],
                  [14.0, 46.0, 37.0, 46.0, 46.0, 0.0, 37.0, 14.0],
                  [37.0, 46.0, 1.0, 46.0, 46.0, 37.0, 0.0, 37.0],
                  [1.0, 46.0, 37.0, 46.0, 46.0, 14.0, 37.0, 0.0]])
    left, right = gtref.nj_tree(D)
    names = list("ABCDEFGH")
    assert gtref.to_newick(left, right, names) == "(((A,H),(((B,E),D),(C,G))),F);"
    order = [names[v] if v < 8 else "" for v in gtref.postorder(left, right, 8)]
    assert order == ["A", "H", "", "B", "E", "", "D", "", "C", "G", "", "", "", "F", ""]


def test_tree_small():
    assert gtref.nj_tree(np.zeros((1, 1))) == -1                       # the reference indexes nodes[1]
    assert gtref.nj_tree(np.array([[0.0, 0.3], [0.3, 0.0]])) == ([0], [1])
    left, right = gtref.nj_tree(np.array([[0.0, 3.0, 5.0], [3.0, 0.0, 4.0], [5.0, 4.0, 0.0]]))
    assert gtref.to_newick(left, right, list("ABC")) == "((A,B),C);"    # every Q off the diagonal is -12: first in row-major order
