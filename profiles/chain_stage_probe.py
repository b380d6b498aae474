"""Runs K4 (chain fill) alone on the anchors of one synthetic genome pair (the benchmark's input shape); the command
profiled by ncu for profiles/r01_k4_*.  usage: chain_stage_probe.py [genome_len]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import chainref  # noqa: E402
from oracle import refmm2  # noqa: E402
from pangraph_b200 import abi, synth  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
anc = synth.ancestor(L, 42)
seqs = [synth.mutate(anc, 20260).tobytes(), synth.mutate(anc, 20261).tobytes()]
idx = abi.Index(seqs, ["0", "1"], "asm10", None, 90)
a = idx.collect_seeds()[0][0]
idx.close()
chainref.ref_sort(refmm2.load_ref(), a)
for rep in range(3):
    t = time.perf_counter()
    u, kept, fpv, seg = abi.chain_rmq(a, 10000, 1000, 1000, 25, 100000, 3, 40, np.float32(0.8 * 0.01 * 19), 0.0)
    print(f"rep {rep}: {len(a)} anchors, {seg[0]} segments ({seg[1]} to host), {len(u)} chains, {1e3 * (time.perf_counter() - t):.1f} ms")
