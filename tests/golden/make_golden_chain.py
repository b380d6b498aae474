"""Regenerates tests/golden/golden_chain.npz from the UNMODIFIED reference (oracle/_ref/libmm2ref.so): sorted anchor sets
(radix_sort_128x order) and what mg_lchain_rmq (lchain.c:250-368) returns for them with pangraph's asm parameters.
Run where /root/reference is mounted:  python tests/golden/make_golden_chain.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import chainref  # noqa: E402
from oracle import refmm2  # noqa: E402

PARAMS = dict(max_dist=10000, inner=1000, bw=1000, skip=25, cap=100000, min_cnt=3, min_sc=40, pen_gap=float(np.float32(0.8 * 0.01 * 19)), pen_skip=0.0)


def cases():
    rng = np.random.default_rng(2026)
    yield "noisy", chainref.synth_anchors(rng, 2400)
    yield "colinear", np.concatenate([chainref.colinear_anchors(rng, 4000), chainref.colinear_anchors(rng, 1500, strand=1, y0=9000)])
    xs = np.repeat(np.arange(1000, 1000 + 3 * 60, 3, dtype=np.uint64), 25)
    ys = np.tile(np.arange(500, 500 + 7 * 25, 7, dtype=np.uint64), 60)
    yield "repeat_array", np.stack([xs, (np.uint64(19) << np.uint64(32)) | ys], axis=1).copy()


if __name__ == "__main__":
    ref = refmm2.load_ref()
    out = {"params": np.array([PARAMS[k] for k in ("max_dist", "inner", "bw", "skip", "cap", "min_cnt", "min_sc", "pen_gap", "pen_skip")])}
    for name, a in cases():
        chainref.ref_sort(ref, a)
        u, kept = chainref.ref_chain(ref, a, PARAMS["max_dist"], PARAMS["inner"], PARAMS["bw"], PARAMS["skip"], PARAMS["cap"],
                                     PARAMS["min_cnt"], PARAMS["min_sc"], np.float32(PARAMS["pen_gap"]), PARAMS["pen_skip"])
        out[name + "_in"], out[name + "_u"], out[name + "_kept"] = a, u, kept
        print(name, len(a), "anchors ->", len(u), "chains,", len(kept), "kept")
    np.savez_compressed(os.path.join(HERE, "golden_chain.npz"), **out)
