"""Second look at gpurun_out/cta_trace.npy (cta_trace_run.py): how fast the long single-CTA problems run under load
(us per anti-diagonal), and how many long CTAs share an SM.  usage: cta_trace_analyze.py [trace.npy]"""
import sys

import numpy as np

rec = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cta_trace.npy")
names = {1: "K5a", 2: "K5b first", 3: "K5b exact", 4: "K5 generic", 5: "K4"}
for k in (2, 3, 4):
    r = rec[(rec["kernel"] == k) & (rec["aux"] >= 500)]
    if len(r) == 0:
        continue
    us = (r["t1"] - r["t0"]) / 1e3 / r["aux"]
    d = (r["t2"] - r["t0"]) / 1e6
    print(f"{names[k]:11s} CTAs with >= 500 rows: {len(r):6d}; us per row p10 {np.percentile(us, 10):.2f} p50 {np.median(us):.2f} p90 {np.percentile(us, 90):.2f} "
          f"p99 {np.percentile(us, 99):.2f}; rows p50 {np.median(r['aux']):.0f} p90 {np.percentile(r['aux'], 90):.0f} max {r['aux'].max()}; "
          f"ms p50 {np.median(d):.2f} p90 {np.percentile(d, 90):.2f} p99 {np.percentile(d, 99):.2f}; tail (t2-t1) ms p50 {np.median((r['t2'] - r['t1']) / 1e6):.3f} p99 {np.percentile((r['t2'] - r['t1']) / 1e6, 99):.3f}")
# long CTAs (>= 2 ms) resident per SM over time
lng = rec[((rec["kernel"] == 2) | (rec["kernel"] == 3) | (rec["kernel"] == 4)) & (rec["t2"] - rec["t0"] >= 2_000_000)]
if len(lng):
    tmin, tmax = rec["t0"].min(), rec["t2"].max()
    nb = int((tmax - tmin) // 1_000_000) + 2
    occ = np.zeros((nb, 160), dtype=np.int16)
    for a, b, s in zip(((lng["t0"] - tmin) // 1_000_000).astype(int), ((lng["t2"] - tmin) // 1_000_000).astype(int), lng["smid"]):
        occ[a:b + 1, s] += 1
    tot = occ.sum(axis=1)
    print(f"long CTAs (>= 2 ms): {len(lng)}; resident at a time: mean {tot.mean():.1f} p90 {np.percentile(tot, 90):.0f} max {tot.max()}; "
          f"SMs holding one: mean {(occ > 0).sum(axis=1).mean():.1f}; SM-ms with 2+ long CTAs: {(occ >= 2).sum()} of {(occ >= 1).sum()}")
# K5a CTA durations by time decile (does the small-fill kernel slow down when long CTAs are around?)
r = rec[rec["kernel"] == 1]
if len(r):
    d = (r["t2"] - r["t0"]) / 1e6
    print(f"K5a CTAs {len(r)}: ms p10 {np.percentile(d, 10):.3f} p50 {np.median(d):.3f} p90 {np.percentile(d, 90):.3f} p99 {np.percentile(d, 99):.3f}")
