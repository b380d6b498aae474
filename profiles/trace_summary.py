"""Summarises the PGMM_TRACE lines of a bench run: DP-service batches per lane (rounds merged, jobs, run / kernel ms) and
how long rounds waited for their waves.  usage: trace_summary.py trace.log"""
import re
import statistics as st
import sys

batch = {}
waits = []
rb = re.compile(r"dp batch \((\w+) lane\): (\d+) rounds, (\d+) jobs, merge ([\d.]+) ms, run ([\d.]+) ms \(kernels ([\d.]+) ms, (\d+) launches\)")
rw = re.compile(r"dp wave: (\d+) short, (\d+) medium, (\d+) long jobs, waited ([\d.]+) ms")
spans = []
rs = re.compile(r"dp span: (\d+) jobs, (\d+) launches; host: prep ([\d.]+) ms, enqueue\+join ([\d.]+) ms, results ([\d.]+) ms; gpu: main stream reached the wave -> first CTA ([-\d.]+) ms, "
                r"first CTA -> last CTA end ([-\d.]+) ms, last CTA end -> main stream after the join ([-\d.]+) ms")
for line in open(sys.argv[1], errors="replace"):
    m = rs.search(line)
    if m:
        spans.append(tuple(float(x) for x in m.groups()))
        continue
    m = rb.search(line)
    if m:
        batch.setdefault(m.group(1), []).append(tuple(float(x) for x in m.groups()[1:]))
        continue
    m = rw.search(line)
    if m:
        waits.append(tuple(float(x) for x in m.groups()))


def q(v, p):
    v = sorted(v)
    return v[min(len(v) - 1, int(p * len(v)))] if v else 0.0


for lane, rows in batch.items():
    r, j, mg, run, k, ln = zip(*rows)
    print(f"{lane:7s} batches {len(rows):6d}  rounds/batch mean {st.mean(r):5.1f} max {max(r):3.0f}  jobs/batch mean {st.mean(j):8.0f}  merge ms mean {st.mean(mg):5.1f}  "
          f"run ms mean {st.mean(run):6.1f} p50 {q(run, .5):6.1f} p90 {q(run, .9):6.1f} max {max(run):6.1f}  kernels ms mean {st.mean(k):6.1f} p90 {q(k, .9):6.1f}  launches mean {st.mean(ln):4.1f}")
if waits:
    w = [x[3] for x in waits]
    print(f"waves {len(w)}: waited ms mean {st.mean(w):6.1f} p50 {q(w, .5):6.1f} p90 {q(w, .9):6.1f} max {max(w):6.1f}")
    for name, sel in (("with long jobs", [x[3] for x in waits if x[2] > 0]), ("medium, no long", [x[3] for x in waits if x[2] == 0 and x[1] > 0]),
                      ("short only", [x[3] for x in waits if x[2] == 0 and x[1] == 0])):
        if sel:
            print(f"  {name:16s} {len(sel):6d} waves: mean {st.mean(sel):6.1f} p50 {q(sel, .5):6.1f} p90 {q(sel, .9):6.1f}")

for name, sel in (("spans, > 1000 jobs (short lane)", [x for x in spans if x[0] > 1000]), ("spans, <= 1000 jobs", [x for x in spans if x[0] <= 1000])):
    if sel:
        cols = list(zip(*sel))
        print(f"{name}: {len(sel)}; host prep mean {st.mean(cols[2]):6.1f} p50 {q(cols[2], .5):6.1f} p90 {q(cols[2], .9):6.1f}; enqueue+join mean {st.mean(cols[3]):6.1f} p50 {q(cols[3], .5):6.1f} p90 {q(cols[3], .9):6.1f}; "
              f"results mean {st.mean(cols[4]):6.1f} p50 {q(cols[4], .5):6.1f} p90 {q(cols[4], .9):6.1f}; "
              f"gpu reach->first CTA mean {st.mean(cols[5]):6.2f} p50 {q(cols[5], .5):6.2f} p90 {q(cols[5], .9):6.2f}; first->last CTA end mean {st.mean(cols[6]):6.1f} p50 {q(cols[6], .5):6.1f} p90 {q(cols[6], .9):6.1f}; "
              f"last end->after join mean {st.mean(cols[7]):6.2f} p50 {q(cols[7], .5):6.2f} p90 {q(cols[7], .9):6.2f}")
