"""One long unbanded fill (K5b) through the stage entry point; the command profiled by ncu for the K5b summary.
usage: k5b_probe.py [qlen] [tlen]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import kswref  # noqa: E402
from pangraph_b200 import abi  # noqa: E402

ql = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
tl = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
rng = np.random.default_rng(3)
q, t = kswref.random_pair(rng, ql, tl, div=0.3, indel=0.02)
for rep in range(3):
    t0 = time.perf_counter()
    ez, cigs, ms = abi.ksw_extd2_batch([len(q)], [len(t)], [0], [0], q, t, [150001], [200], [-1], [kswref.FLAG_FILL1], 1, 9, 1, 16, 2, 41, 1)
    print(f"rep {rep}: {len(q)} x {len(t)}: kernel {ms:.2f} ms, wall {1e3 * (time.perf_counter() - t0):.1f} ms, score {int(ez[0, 8])}, {len(cigs[0])} cigar ops")
