"""Fixture generator (run where /root/reference is mounted): the first two records of the reference's bundled
data/ecoli.fa.gz (BASELINE config 2) and data/klebs.fa.gz (config 3), upper-cased, packed 2 bits per base with the few
non-ACGT letters kept as (position, letter) exceptions -> tests/golden/real_pairs.npz.  Inputs only: the expected hits
are computed at test time by the reference's own C (oracle/_ref), which travels to the GPU box."""
import gzip
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def first_records(path, n):
    recs = []
    with gzip.open(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                if len(recs) >= n:
                    break
                recs.append([line[1:].split()[0], []])
            else:
                recs[-1][1].append(line.strip().upper())
    return [(name, "".join(parts)) for name, parts in recs]


def pack(seq):
    a = np.frombuffer(seq.encode(), dtype=np.uint8)
    code = np.full(256, 255, dtype=np.uint8)
    code[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
    c = code[a]
    exc_pos = np.flatnonzero(c == 255).astype(np.int64)
    exc_chr = a[exc_pos].copy()
    c[exc_pos] = 0
    pad = (-len(c)) % 4
    c = np.concatenate([c, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    packed = (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)
    return packed, np.int64(len(a)), exc_pos, exc_chr


def main():
    out = {}
    for tag in ("ecoli", "klebs"):
        for i, (name, seq) in enumerate(first_records(f"/root/reference/data/{tag}.fa.gz", 2)):
            packed, n, ep, ec = pack(seq)
            out[f"{tag}{i}_packed"], out[f"{tag}{i}_len"], out[f"{tag}{i}_exc_pos"], out[f"{tag}{i}_exc_chr"] = packed, n, ep, ec
            out[f"{tag}{i}_name"] = np.array(name)
    np.savez(os.path.join(HERE, "real_pairs.npz"), **out)


if __name__ == "__main__":
    main()
