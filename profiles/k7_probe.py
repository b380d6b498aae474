"""K7 (mash distance) timing probe: n genomes of a synthetic family through pgmm_mash_distance, stage times from CUDA events
inside the library, the oracle restatement of the reference's serial path timed beside it on the host.
usage: python profiles/k7_probe.py [n] [length] -> one JSON line (also appended to gpurun_out/k7_probe.json)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import gtref
from pangraph_b200 import guide_tree as gt, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
length = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
gs = [g for _, g in synth.genomes(n, length=length)]
gt.mash_distance(gs[:2])  # context, module load
best = None
for _ in range(3):
    t = time.perf_counter()
    D, st = gt.mash_distance(gs, with_stats=True)
    st["wall_ms"] = (time.perf_counter() - t) * 1e3
    if best is None or st["wall_ms"] < best["wall_ms"]:
        best = st
t = time.perf_counter()
want = gtref.mash_distance(gs)
cpu_ms = (time.perf_counter() - t) * 1e3
t = time.perf_counter()
tree = gt.neighbor_joining(D)
nj_ms = (time.perf_counter() - t) * 1e3
bases = best["bases"]
line = dict(n=n, length=length, identical=bool(np.array_equal(D, want)), **best, cpu_oracle_ms=cpu_ms, nj_host_ms=nj_ms,
            sketch_gbp_per_s=bases / best["sketch_ms"] / 1e6, e2e_gbp_per_s=bases / best["wall_ms"] / 1e6,
            cpu_gbp_per_s=bases / cpu_ms / 1e6)
print(json.dumps(line))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "k7_probe.json"), "a") as f:
    f.write(json.dumps(line) + "\n")
