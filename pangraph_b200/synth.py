"""Deterministic synthetic genomes for the benchmark and the parity tests (SURVEY 8d / BASELINE.md section 3).

ancestor: iid uniform ACGT from PCG64(seed).  genome i (seed base+i): substitutions at `sub` per site, single-base
insertions and deletions at `indel` each, then `n_rearr` structural rearrangements drawn in the fixed proportion
4 inversions : 3 translocations : 2 deletions : 1 novel insertion, lengths log-uniform in [len_lo, len_hi].
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[_ACGT] = np.frombuffer(b"TGCA", dtype=np.uint8)


def ancestor(length, seed=42):
    rng = np.random.Generator(np.random.PCG64(seed))
    return _ACGT[rng.integers(0, 4, size=length)]


def revcomp(a):
    return _COMP[a[::-1]]


def mutate(anc, seed, sub=0.005, indel=0.0005, n_rearr=10, len_lo=1000, len_hi=50000):
    rng = np.random.Generator(np.random.PCG64(seed))
    g = anc.copy()
    n = len(g)
    # substitutions
    pos = np.flatnonzero(rng.random(n) < sub)
    idx = np.searchsorted(_ACGT, g[pos])
    g[pos] = _ACGT[(idx + rng.integers(1, 4, size=len(pos))) % 4]
    # 1-bp deletions and insertions
    keep = rng.random(n) >= indel
    ins = rng.random(n) < indel
    out_len = keep.astype(np.int64) + ins.astype(np.int64)
    dst = np.cumsum(out_len) - out_len
    out = np.empty(int(out_len.sum()), dtype=np.uint8)
    out[dst[keep] + ins[keep]] = g[keep]
    out[dst[ins]] = _ACGT[rng.integers(0, 4, size=int(ins.sum()))]
    g = out
    # structural rearrangements
    kinds = (["inv"] * 4 + ["trans"] * 3 + ["del"] * 2 + ["ins"])
    lo, hi = np.log(len_lo), np.log(min(len_hi, max(len_lo + 1, len(g) // 8)))
    for r in range(n_rearr):
        kind = kinds[r % len(kinds)]
        L = int(np.exp(rng.uniform(lo, hi)))
        p = int(rng.integers(0, max(1, len(g) - L)))
        if kind == "inv":
            g = np.concatenate([g[:p], revcomp(g[p:p + L]), g[p + L:]])
        elif kind == "trans":
            seg = g[p:p + L]
            rest = np.concatenate([g[:p], g[p + L:]])
            t = int(rng.integers(0, len(rest)))
            g = np.concatenate([rest[:t], seg, rest[t:]])
        elif kind == "del":
            g = np.concatenate([g[:p], g[p + L:]])
        else:
            g = np.concatenate([g[:p], _ACGT[rng.integers(0, 4, size=L)], g[p:]])
    return g


def genomes(n, length=5_000_000, anc_seed=42, base_seed=20260, **kw):
    """n genomes as ASCII bytes objects, names g0000.. ."""
    anc = ancestor(length, anc_seed)
    return [(f"g{i:04d}", mutate(anc, base_seed + i, **kw).tobytes()) for i in range(n)]
