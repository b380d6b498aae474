"""Runs K4 (chain fill) alone on a genome-like anchor set; the command profiled by ncu for profiles/r01_k4_*."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import chainref  # noqa: E402
from pangraph_b200 import abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 190000
rng = np.random.default_rng(5)
a = chainref.colinear_anchors(rng, n)
a = a[np.argsort(a[:, 0], kind="stable")]
for rep in range(3):
    t = time.perf_counter()
    u, kept, fpv, seg = abi.chain_rmq(a, 10000, 1000, 1000, 25, 100000, 3, 40, np.float32(0.152), 0.0)
    print(f"rep {rep}: {len(a)} anchors, {seg[0]} segments ({seg[1]} to host), {len(u)} chains, {1e3 * (time.perf_counter() - t):.1f} ms")
