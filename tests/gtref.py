"""ctypes helpers for the guide-tree restatement in oracle/liboracle.so (oracle/guide_tree_oracle.c)."""
import ctypes as C

import numpy as np

import kswref


def _lib():
    lib = kswref.load_oracle()
    lib.orc_mash_hash.restype = C.c_uint64
    lib.orc_mash_hash.argtypes = [C.c_uint64, C.c_uint64]
    lib.orc_mash_sketch.restype = C.c_int64
    return lib


def mash_hash(x, mask):
    return int(_lib().orc_mash_hash(x, mask))


def mash_sketch(seq, sid, k=15, w=100):
    """-> [(value, position)] exactly as minimizers_sketch returns them (order, repetitions); [] is its error case"""
    s = seq.encode() if isinstance(seq, str) else bytes(seq)
    lib = _lib()
    cap = max(16, len(s) // 4)
    while True:
        val, pos = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        n = lib.orc_mash_sketch(s, C.c_int64(len(s)), C.c_uint64(sid), k, w, C.c_void_p(val.ctypes.data), C.c_void_p(pos.ctypes.data),
                                C.c_int64(cap))
        if n < 0:
            raise ValueError("k must be < 32 and w < 256")
        if n <= cap:
            return [(int(val[i]), int(pos[i])) for i in range(n)]
        cap = int(n)


def mash_distance(seqs, k=15, w=100):
    """-> n x n float64, or an int error code (-(1+i): sequence i has no minimizer; -1000000: no sequences)"""
    lib = _lib()
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    n = len(bs)
    arr = (C.c_char_p * max(n, 1))(*bs)
    lens = np.array([len(b) for b in bs] or [0], np.int64)
    out = np.zeros((max(n, 1), max(n, 1)), np.float64)
    rc = lib.orc_mash_distance(arr, C.c_void_p(lens.ctypes.data), n, k, w, C.c_void_p(out.ctypes.data))
    return out if rc == 0 else int(rc)


def nj_q_matrix(D):
    D = np.ascontiguousarray(D, np.float64)
    Q = np.zeros_like(D)
    _lib().orc_nj_q_matrix(C.c_void_p(D.ctypes.data), D.shape[0], C.c_void_p(Q.ctypes.data))
    return Q


def nj_dist(D, i, j):
    D = np.ascontiguousarray(D, np.float64)
    dn = np.zeros(D.shape[0], np.float64)
    _lib().orc_nj_dist(C.c_void_p(D.ctypes.data), D.shape[0], int(i), int(j), C.c_void_p(dn.ctypes.data))
    return dn


def nj_tree(D):
    """-> (left, right): children of the nodes n, n+1, ..., 2n-2 (the root); or an int error code"""
    D = np.ascontiguousarray(D, np.float64)
    n = D.shape[0]
    left, right = np.zeros(max(n - 1, 1), np.int32), np.zeros(max(n - 1, 1), np.int32)
    rc = _lib().orc_nj_tree(C.c_void_p(D.ctypes.data), n, C.c_void_p(left.ctypes.data), C.c_void_p(right.ctypes.data))
    return (left[:n - 1].tolist(), right[:n - 1].tolist()) if rc == 0 else int(rc)


def to_newick(left, right, names):
    """Clade::to_newick (PG/tree/newick.rs:11-38) over the (left, right) arrays: internal nodes carry no label"""
    n = len(names)

    def rec(v):
        return names[v] if v < n else "(" + rec(left[v - n]) + "," + rec(right[v - n]) + ")"
    return rec(2 * n - 2) + ";"


def postorder(left, right, n):
    """clade.rs:49-71: left subtree, right subtree, the node"""
    out = []

    def rec(v):
        if v >= n:
            rec(left[v - n]), rec(right[v - n])
        out.append(v)
    rec(2 * n - 2)
    return out
