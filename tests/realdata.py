"""Loader of tests/golden/real_pairs.npz (made by tests/golden/make_real_pairs.py from the reference's bundled data)."""
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_pairs.npz")


def load_pair(tag):
    """-> ([seq0, seq1] as ASCII bytes, [record names]) of the first two genomes of data/<tag>.fa.gz"""
    z = np.load(PATH)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs, names = [], []
    for i in range(2):
        p = z[f"{tag}{i}_packed"]
        n = int(z[f"{tag}{i}_len"])
        c = np.stack([p & 3, (p >> 2) & 3, (p >> 4) & 3, (p >> 6) & 3], axis=1).reshape(-1)[:n]
        a = acgt[c]
        a[z[f"{tag}{i}_exc_pos"]] = z[f"{tag}{i}_exc_chr"]
        seqs.append(a.tobytes())
        names.append(str(z[f"{tag}{i}_name"]))
    return seqs, names
