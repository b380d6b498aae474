"""Runs the benchmark's resident-input rounds with the CTA trace on and summarises what overlapped on the GPU.
usage: cta_trace_run.py [workers] [rounds]   -> gpurun_out/cta_trace.npy + a summary on stdout"""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, ".")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 36
N = int(sys.argv[2]) if len(sys.argv) > 2 else 108
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("PGMM_CONTEXTS", str(max(8, W)))
os.environ.setdefault("PGMM_ARENA_GB", "2")
from pangraph_b200 import abi, synth  # noqa: E402

L = abi.lib()
n_pool = min(N, max(W, 12))
anc = synth.ancestor(5_000_000, 42)
pairs = [([synth.mutate(anc, 20260 + 2 * p).tobytes(), synth.mutate(anc, 20261 + 2 * p).tobytes()], [str(2 * p), str(2 * p + 1)]) for p in range(n_pool)]


def one(p):
    idx = abi.Index(*pairs[p % n_pool], "asm10", None, 90, resident_only=True)
    idx.build()
    out = idx.map_self(raw=True)
    idx.close()
    return out


pool = ThreadPoolExecutor(W)
list(pool.map(one, range(2 * W)))  # warm-up
cap = 4_000_000
L.pgmm_cta_trace_begin.argtypes = [C.c_uint64]
L.pgmm_cta_trace_end.argtypes = [C.c_void_p, C.c_uint64]
L.pgmm_cta_trace_end.restype = C.c_int64
assert L.pgmm_cta_trace_begin(cap) == 0
t0 = time.perf_counter()
list(pool.map(one, range(N)))
wall = time.perf_counter() - t0
rec = np.zeros(cap, dtype=[("t0", "u8"), ("t1", "u8"), ("t2", "u8"), ("kernel", "u4"), ("block", "u4"), ("smid", "u4"), ("aux", "u4")])
n = L.pgmm_cta_trace_end(rec.ctypes.data, cap)
rec = rec[:min(n, cap)]
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/cta_trace.npy", rec)
print(f"{N} rounds, {W} in flight: {wall:.2f} s wall ({N / wall:.1f} rounds/s), {n} CTA records")
names = {1: "K5a", 2: "K5b first", 3: "K5b exact", 4: "K5 generic", 5: "K4"}
span = (rec["t2"].max() - rec["t0"].min()) / 1e9
print(f"traced span {span:.2f} s")
for k, nm in names.items():
    r = rec[rec["kernel"] == k]
    if len(r) == 0:
        continue
    d = (r["t2"] - r["t0"]) / 1e6
    fill = (r["t1"] - r["t0"]) / 1e6
    print(f"{nm:11s} CTAs {len(r):8d}  sum {d.sum() / 1e3:8.2f} SM-s  ({d.sum() / 1e3 / span:6.1f} CTAs resident on average)  per CTA ms: mean {d.mean():.3f} "
          f"p50 {np.median(d):.3f} p99 {np.percentile(d, 99):.3f} max {d.max():.3f}; main loop share {fill.sum() / max(d.sum(), 1e-9):.2f}")
# how many SMs hold at least one traced CTA over time (1 ms buckets)
tmin = rec["t0"].min()
nb = int(span * 1e3) + 2
busy = np.zeros((nb, 160), dtype=bool)
for k in (2, 3, 4, 5, 1):
    r = rec[rec["kernel"] == k]
    b0 = ((r["t0"] - tmin) // 1_000_000).astype(int)
    b1 = ((r["t2"] - tmin) // 1_000_000).astype(int)
    for a, b, s in zip(b0, b1, r["smid"]):
        busy[a:b + 1, s] = True
print("SMs with a traced CTA resident, averaged over 1 ms buckets: %.1f of %d" % (busy.sum(axis=1).mean(), int(rec["smid"].max()) + 1))
