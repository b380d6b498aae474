"""One 2 x 5 Mbp round ALONE (nothing else in flight) with PGMM_TRACE=1: the per-phase wall and CPU times of a round and the
trace lines of its DP waves / chain fill.  usage: single_round_trace.py [reps]   (set the env yourself: PGMM_TRACE=1)"""
import os
import sys
import time

sys.path.insert(0, ".")
import bench  # noqa: E402
from pangraph_b200 import abi  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
(seqs, names), = bench.make_pairs(1, 0, 5_000_000)
for rep in range(reps):
    abi.get_stats(reset=True)
    c0, t0 = os.times(), time.perf_counter()
    idx = abi.Index(seqs, names, "asm10", None, 90, resident_only=True)
    idx.build()
    got = idx.map_self()
    idx.close()
    c1, t1 = os.times(), time.perf_counter()
    st = abi.get_stats()
    print(f"rep {rep}: {sum(len(g) for g in got)} hits, wall {1e3 * (t1 - t0):.1f} ms, cpu {1e3 * (c1.user + c1.system - c0.user - c0.system):.1f} ms; "
          + ", ".join(f"{k} {st[k]:.1f}" for k in ("index_ms", "t_encode", "t_seed", "t_chain_sort", "t_chain_fill", "t_chain_rest", "t_dp", "t_stitch",
                                                  "t_final", "chain_kernel_ms", "chain_iterations", "chain_launches", "dp_waves", "launches")), flush=True)
