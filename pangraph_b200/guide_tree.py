"""The guide tree of `pangraph build` through the C-ABI (include/pgmm_b200.h, Part 5): mash distance on the GPU (K7),
neighbour joining / Newick / balancing on the host (C++).  Mirrors PG = packages/pangraph/src:
PG/distance/mash/mash_distance.rs (mash_distance), PG/tree/neighbor_joining.rs (build_tree_using_neighbor_joining),
PG/tree/newick.rs (parse_newick, to_newick, build_tree_from_newick), PG/tree/balance.rs (balance), PG/tree/clade.rs (postorder).

A tree over n leaves is a GuideTree(n, left, right, names): leaves 0..n-1, internal node n + t has children left[t], right[t],
the root is 2n - 2; `children()` is the form sharding.TreeSchedule / sharding.run_tree take."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi

MASH_K, MASH_W = 15, 100  # MinimizersParams::default (PG/distance/mash/minimizer.rs:8-14)


class GuideTreeError(RuntimeError):
    pass


@dataclass
class GuideTree:
    n: int
    left: list
    right: list
    names: list = field(default_factory=list)

    @property
    def root(self):
        return 0 if self.n == 1 else 2 * self.n - 2

    def children(self):
        """children[node] = (left, right) or None for a leaf -- the input of sharding.TreeSchedule"""
        return [None] * self.n + list(zip(self.left, self.right))

    def postorder(self):
        L = abi.lib()
        order = (C.c_int32 * (2 * self.n - 1))()
        m = L.pgmm_tree_postorder(self.n, _i32(self.left), _i32(self.right), order)
        return [int(order[i]) for i in range(m)]

    def to_newick(self, names=None):
        names = self.names if names is None else names
        assert len(names) == self.n, "one name per leaf"
        L = abi.lib()
        L.pgmm_newick_write.restype = C.c_void_p
        arr = (C.c_char_p * self.n)(*[s.encode() if isinstance(s, str) else s for s in names])
        p = L.pgmm_newick_write(self.n, arr, _i32(self.left), _i32(self.right))
        try:
            return C.string_at(p).decode()
        finally:
            _free(p)

    def balance(self):
        L = abi.lib()
        ol, orr = (C.c_int32 * max(1, self.n - 1))(), (C.c_int32 * max(1, self.n - 1))()
        rc = L.pgmm_tree_balance(self.n, _i32(self.left), _i32(self.right), ol, orr)
        if rc != 0:
            raise GuideTreeError(f"pgmm_tree_balance -> {rc}")
        return GuideTree(self.n, [int(ol[i]) for i in range(self.n - 1)], [int(orr[i]) for i in range(self.n - 1)], list(self.names))


def _i32(xs):
    return (C.c_int32 * max(1, len(xs)))(*xs)


def _free(p):
    L = abi.lib()
    L.pgmm_free.argtypes = [C.c_void_p]
    L.pgmm_free.restype = None
    L.pgmm_free(p)


def mash_distance(seqs, k=MASH_K, w=MASH_W, with_stats=False):
    """mash_distance(graphs, params) for singleton graphs: seqs[i] = the sequence of graph i (str or bytes).
    -> n x n float64.  Raises where the reference panics (a sequence without minimizers, k >= 32, w >= 256, no sequences)."""
    L = abi.lib()
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    n = len(bs)
    if n == 0:
        raise GuideTreeError("mash distance of no sequences (the reference's array![[]] fails neighbour joining's shape assert)")
    arr = (C.c_char_p * n)(*bs)
    lens = (C.c_int64 * n)(*[len(b) for b in bs])
    out = np.zeros((n, n), np.float64)
    st = (C.c_double * 10)()
    L.pgmm_mash_distance.restype = C.c_int
    rc = L.pgmm_mash_distance(n, arr, lens, k, w, C.c_void_p(out.ctypes.data), st)
    if rc > 0:
        raise GuideTreeError(f"no minimizer found for sequence {rc - 1} during mash distance evaluation")
    if rc < 0:
        raise GuideTreeError({-1: "k must be in 1..31 and w in 1..255", -2: "no sequences", -3: "2k + bits(n) exceeds 64",
                              -4: "a sequence of 2^31 bases or more", -5: "the incidence bitmap does not fit the device"}.get(rc, str(rc)))
    if with_stats:
        keys = ("upload_ms", "sketch_ms", "sort_ms", "pair_ms", "bases", "tiles", "minimizers", "unique_keys", "shared_values", "launches")
        return out, dict(zip(keys, (float(v) for v in st)))
    return out


def neighbor_joining(dist, names=None):
    """build_tree_using_neighbor_joining on a distance matrix -> GuideTree"""
    L = abi.lib()
    D = np.ascontiguousarray(dist, np.float64)
    n = D.shape[0]
    assert D.shape == (n, n)
    left, right = (C.c_int32 * max(1, n - 1))(), (C.c_int32 * max(1, n - 1))()
    L.pgmm_nj_tree.restype = C.c_int
    rc = L.pgmm_nj_tree(n, C.c_void_p(D.ctypes.data), left, right)
    if rc != 0:
        raise GuideTreeError({-1: "neighbour joining needs two sequences or more", -2: "NaN in the distance matrix"}.get(rc, str(rc)))
    return GuideTree(n, [int(left[i]) for i in range(n - 1)], [int(right[i]) for i in range(n - 1)], list(names) if names else [])


def build_tree_using_neighbor_joining(seqs, names=None):
    """PG/tree/neighbor_joining.rs:16-35: mash distances (k = 15, w = 100) on the GPU, then the joins"""
    return neighbor_joining(mash_distance(seqs), names)


def parse_newick(text):
    """PG/tree/newick.rs:43-62 -> GuideTree with the leaf labels in order of appearance; GuideTreeError carries the
    reference's message"""
    L = abi.lib()
    n, nb = C.c_int32(0), C.c_int64(0)
    names, left, right = C.c_void_p(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
    err = C.create_string_buffer(1024)
    L.pgmm_newick_parse.restype = C.c_int
    rc = L.pgmm_newick_parse(text.encode() if isinstance(text, str) else text, C.byref(n), C.byref(names), C.byref(nb), C.byref(left),
                             C.byref(right), err, len(err))
    if rc != 0:
        raise GuideTreeError(err.value.decode(errors="replace"))
    try:
        raw = C.string_at(names, nb.value)
        labels = [s.decode() for s in raw.split(b"\0")[:n.value]]
        return GuideTree(n.value, [int(left[i]) for i in range(n.value - 1)], [int(right[i]) for i in range(n.value - 1)], labels)
    finally:
        _free(names), _free(C.cast(left, C.c_void_p)), _free(C.cast(right, C.c_void_p))


def build_tree_from_newick(text, seq_names):
    """PG/tree/newick.rs:69-141: the tree of a Newick text with its leaves matched to the FASTA names -> GuideTree whose leaf
    i is record i of `seq_names`.  Same checks and messages as the reference."""
    t = parse_newick(text)
    by_name = {}
    for i, name in enumerate(seq_names):
        if name in by_name:
            raise GuideTreeError(f"Duplicate FASTA sequence name '{name}'")
        by_name[name] = i
    leaf_of = []
    for label in t.names:  # leaves are numbered in the order attach_graphs meets them (left to right)
        if label not in by_name:
            raise GuideTreeError(f"Newick leaf '{label}' has no matching FASTA record")
        leaf_of.append(by_name.pop(label))
    if by_name:
        raise GuideTreeError(f"FASTA records [{', '.join(sorted(by_name))}] are not present in the guide tree")
    m = t.n
    ren = lambda v: leaf_of[v] if v < m else v
    return GuideTree(m, [ren(v) for v in t.left], [ren(v) for v in t.right], list(seq_names))


def main(argv=None):
    """`python -m pangraph_b200.guide_tree a.fa [b.fa.gz ...]`: the guide tree `pangraph build` would use for these genomes
    (FASTA in as the reference reads it, mash distance on the GPU, neighbour joining), printed as the Newick line the
    reference logs (`Guide tree (newick): ...`, commands/build/build_run.rs:102); `--balance` prints the bisected tree too."""
    import argparse
    import sys
    from . import fasta
    ap = argparse.ArgumentParser(prog="python -m pangraph_b200.guide_tree", description=main.__doc__)
    ap.add_argument("fasta", nargs="+")
    ap.add_argument("--balance", action="store_true")
    ap.add_argument("--distances", action="store_true", help="print the mash distance matrix as well")
    args = ap.parse_args(argv)
    recs = fasta.read_many(args.fasta)
    names = [r.seq_name for r in recs]
    D = mash_distance([r.seq for r in recs])
    if args.distances:
        for name, row in zip(names, D):
            print(name, " ".join(f"{v:.6f}" for v in row), file=sys.stderr)
    tree = neighbor_joining(D, names)
    print(tree.to_newick())
    if args.balance:
        print(tree.balance().to_newick())
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
