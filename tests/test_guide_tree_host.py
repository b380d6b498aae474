"""Host side of the guide-tree path through the C-ABI (no GPU): neighbour joining against the oracle restatement and the
reference's vectors (PG/tree/neighbor_joining.rs:111-288), Newick in / out against PG/tree/newick.rs:283-345, balance
(PG/tree/balance.rs:65-110), postorder (PG/tree/clade.rs:97-123), and the tree as input of the merge scheduler."""
import numpy as np
import pytest

import gtref
from pangraph_b200 import guide_tree as gt
from pangraph_b200 import sharding
from test_oracle_guide_tree import WIKI


def test_nj_reference_trees():
    t = gt.neighbor_joining(WIKI, list("ABCDE"))
    assert t.to_newick() == "((((A,B),C),D),E);"
    assert [t.names[v] if v < 5 else "" for v in t.postorder()] == ["A", "B", "", "C", "", "D", "", "E", ""]
    D = np.array([[0.0, 46.0, 37.0, 46.0, 46.0, 14.0, 37.0, 1.0], [46.0, 0.0, 46.0, 7.0, 1.0, 46.0, 46.0, 46.0],
                  [37.0, 46.0, 0.0, 46.0, 46.0, 37.0, 1.0, 37.0], [46.0, 7.0, 46.0, 0.0, 7.0, 46.0, 46.0, 46.0],
                  [46.0, 1.0, 46.0, 7.0, 0.0, 46.0, 46.0, 46.0], [14.0, 46.0, 37.0, 46.0, 46.0, 0.0, 37.0, 14.0],
                  [37.0, 46.0, 1.0, 46.0, 46.0, 37.0, 0.0, 37.0], [1.0, 46.0, 37.0, 46.0, 46.0, 14.0, 37.0, 0.0]])
    t = gt.neighbor_joining(D, list("ABCDEFGH"))
    assert t.to_newick() == "(((A,H),(((B,E),D),(C,G))),F);"
    assert [t.names[v] if v < 8 else "" for v in t.postorder()] == ["A", "H", "", "B", "E", "", "D", "", "C", "G", "", "", "", "F", ""]


@pytest.mark.parametrize("n,seed", [(2, 0), (3, 1), (5, 2), (9, 3), (17, 4), (40, 5), (130, 6)])
def test_nj_matches_oracle_on_mash_like_matrices(n, seed):
    """distances of the form 1 - c / d (what mash_distance produces): near-ties everywhere, so the order of the row and column
    sums and the first-minimum rule decide the joins -- product and oracle must take the same ones"""
    rng = np.random.default_rng(seed)
    own = rng.integers(50, 400, n)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            D[i, j] = D[j, i] = 1.0 - float(rng.integers(0, own[i] + 1)) / float(own[i])
    t = gt.neighbor_joining(D)
    assert (t.left, t.right) == gtref.nj_tree(D)
    # clustered: two families with small distances inside, 1.0 across
    for i in range(n):
        for j in range(i + 1, n):
            D[i, j] = D[j, i] = 1.0 if (i % 2) != (j % 2) else float(rng.integers(1, 9)) / 9.0
    t = gt.neighbor_joining(D)
    assert (t.left, t.right) == gtref.nj_tree(D)


def test_nj_errors():
    with pytest.raises(gt.GuideTreeError):
        gt.neighbor_joining(np.zeros((1, 1)))
    D = WIKI.copy()
    D[1, 2] = D[2, 1] = np.nan
    with pytest.raises(gt.GuideTreeError):
        gt.neighbor_joining(D)
    assert gtref.nj_tree(D) == -2


@pytest.mark.parametrize("text,want", [
    ("((A,B),(C,D));", "((A,B),(C,D));"), ("((A:0.1,B:0.2):0.3,C:0.4);", "((A,B),C);"), ("((A,B)inner,C)root;", "((A,B),C);"),
    ("(\n  (A , B) ,\n  ( C, D )\n);\n", "((A,B),(C,D));"), ("('foo bar',B);", "(foo bar,B);"), ("('it''s',B);", "(it's,B);"),
    ("((A,B),C)", "((A,B),C);"), ("(A:1e-3,B:2.5E+2);", "(A,B);"), ("A;", "A;"), ("(été, B );", "(été,B);")])
def test_newick_round_trip(text, want):
    assert gt.parse_newick(text).to_newick() == want


@pytest.mark.parametrize("text,want", [
    ("", "Newick input is empty"), ("   \n  ", "Newick input is empty"),
    ("((A,B);", "Newick: expected ')' or ',' at position 6, found ';'"),
    ("A,B);", "Newick: unexpected trailing content at position 1: ',B);'"),
    ("(A,B,C);", "Newick: internal node has 3 children; only strictly bifurcating trees are supported"),
    ("(A);", "Newick: internal node has 1 children; only strictly bifurcating trees are supported"),
    ("(,B);", "Newick: leaf without a name at position 1"),
    ("(A,B);xyz", "Newick: unexpected trailing content at position 6: 'xyz'"),
    ("(A:,B);", "Newick: expected a number after ':' at position 3"),
    ("((A,B),C", "Newick: unexpected end of input, expected ')'")])
def test_newick_rejects_malformed_input(text, want):
    with pytest.raises(gt.GuideTreeError) as e:
        gt.parse_newick(text)
    assert str(e.value) == want


def test_build_tree_from_newick():
    t = gt.build_tree_from_newick("((A,B),C);", ["A", "B", "C"])
    assert [v for v in t.postorder() if v < 3] == [0, 1, 2] and t.to_newick() == "((A,B),C);"
    t = gt.build_tree_from_newick("((C,A),B);", ["A", "B", "C"])  # leaves follow the FASTA order, the topology follows the text
    assert (t.left, t.right) == ([2, 3], [0, 1]) and t.to_newick() == "((C,A),B);"
    for text, names, want in (("((A,B),Z);", ["A", "B", "C"], "Newick leaf 'Z' has no matching FASTA record"),
                              ("(A,B);", ["A", "B", "C"], "FASTA records [C] are not present in the guide tree"),
                              ("((A,B),A);", ["A", "B"], "Newick leaf 'A' has no matching FASTA record"),
                              ("(A,B);", ["A", "A"], "Duplicate FASTA sequence name 'A'")):
        with pytest.raises(gt.GuideTreeError) as e:
            gt.build_tree_from_newick(text, names)
        assert str(e.value) == want


def test_deep_caterpillar_newick():
    """a comb over 20 000 leaves nests 20 000 parentheses: no recursion anywhere on the path"""
    n = 20_000
    text = "(" * (n - 1) + "L0" + "".join(f",L{i})" for i in range(1, n)) + ";"
    t = gt.parse_newick(text)
    assert t.n == n and t.to_newick() == text
    assert t.postorder()[:3] == [0, 1, n] and len(t.postorder()) == 2 * n - 1
    b = t.balance()
    assert [v for v in b.postorder() if v < n] == list(range(n))
    depth = {}
    for v in reversed(b.postorder()):
        depth.setdefault(v, 0)
        if v >= n:
            depth[b.left[v - n]] = depth[b.right[v - n]] = depth[v] + 1
    assert max(depth.values()) == 15  # ceil(log2 20000)


def test_balance_reference_vector():
    t = gt.parse_newick("(((A,H),(((B,E),D),(C,G))),F);")
    assert t.balance().to_newick() == "(((A,H),(B,E)),((D,C),(G,F)));"
    assert gt.parse_newick("(((A,B),(C,D)),(G,H));").postorder() == [0, 1, 6, 2, 3, 7, 8, 4, 5, 9, 10]
    one = gt.parse_newick("A;")
    assert one.balance().to_newick() == "A;" and one.postorder() == [0]


def test_tree_feeds_the_merge_scheduler():
    """the joins of neighbour joining as the ready-queue of the merge rounds: one rank replays the reference's post-order,
    several ranks get every ready merge of a wave"""
    t = gt.neighbor_joining(WIKI, list("ABCDE"))
    sched = sharding.TreeSchedule(t.children())
    assert sched.postorder() == t.postorder()
    waves = []
    while not sched.finished():
        wave = sched.next_wave(4)
        waves.append(sorted(v for v, _ in wave))
        sched.complete([v for v, _ in wave])
    assert waves == [[5], [6], [7], [8]]  # a comb: one merge at a time
    b = sharding.TreeSchedule(t.balance().children())
    first = b.next_wave(4)
    assert len(first) == 2 and len({r for _, r in first}) == 2  # balanced: two independent merges on two ranks
