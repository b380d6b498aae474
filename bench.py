#!/usr/bin/env python
"""Benchmark of the alignment hot path of `pangraph build` (BASELINE.json metric: Gbp aligned / s).

One ROUND = one find_matches call of a guide-tree leaf merge: two related synthetic 5-Mbp genomes (1 % divergence,
10 rearrangements each) are indexed and aligned all-vs-all, i.e. mm_idx_str + mm_mapopt_update + one
mm_map per sequence in the reference, index kernels + pgmm_map_batch here.
One STEP = `--rounds-per-step` (default 192) such rounds, `--workers` (default 64) of them in flight at any moment: sibling leaf merges of the guide tree are independent
(merge_graphs only reads its two children), so a rank keeps several of them in flight, one host thread and one CUDA
stream each; their DP waves are merged across rounds by the library's DP service (dp_service.cu).  bp per step = total length of the genomes of its rounds.

  python bench.py [--gpus N --steps K --warmup W]        our CUDA path (N>1: one rank per GPU under torchrun; every
                                                          rank aligns its own pairs -- leaf merges are independent --
                                                          and the match lists are gathered on rank 0 over NCCL)
  python bench.py --impl reference [...]                  the reference's own C (oracle/_ref) on the host cores

`value`  : inputs resident in HBM before the timed region (pgmm_idx_upload done; timed: index kernels + pgmm_map_self)
`e2e`    : the same round through the reference-facing C-ABI with HOST buffers (mm_idx_str, mm_mapopt_update,
           pgmm_map_batch), host->device and device->host copies inside the timed region.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# many independent streams (one per round in flight + one per DP size class): give them their own hardware queues,
# otherwise unrelated kernels serialise behind each other.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "Gbp aligned/sec for `pangraph build` alignment rounds (find_matches)"
UNIT = "Gbp/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_pairs(n_pairs, first_pair, length):
    """Pair p = genomes (2p, 2p+1) of the star family around the PCG64(42) ancestor (seeds 20260+i)."""
    from pangraph_b200 import synth
    anc = synth.ancestor(length, 42)
    pairs = []
    for p in range(first_pair, first_pair + n_pairs):
        a = synth.mutate(anc, 20260 + 2 * p).tobytes()
        b = synth.mutate(anc, 20260 + 2 * p + 1).tobytes()
        pairs.append(([a, b], [str(2 * p), str(2 * p + 1)]))
    return pairs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = max(smax)
        out["reasons"] = sorted(reasons)
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference_step(refmm2, rounds, threads):
    """The reference's path for every round of a step: index, mid_occ, then one mm_map per query; the queries of all
    rounds share one pool of `threads` host threads (ctypes drops the GIL inside the C calls)."""
    from concurrent.futures import ThreadPoolExecutor
    lib = refmm2.load_ref()
    idxs = [refmm2.Index(lib, seqs, names, "asm10", None, 90) for seqs, names in rounds]
    try:
        work = [(ix, i) for ix in idxs for i in range(len(ix.seqs))]
        with ThreadPoolExecutor(max(1, threads)) as ex:
            return list(ex.map(lambda t: len(t[0].map_one(t[1])), work))
    finally:
        for ix in idxs:
            ix.close()


def reference_arm(args):
    """The reference's own CPU implementation (oracle/_ref = its vendored minimap2 C, unmodified) on a bounded sample of
    the same workload: the first `ref_sample_len` bases of each genome of the pair, all host threads it can use
    (one mm_map per sequence: the reference parallelises over queries, align_with_minimap2_lib.rs:64-74)."""
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    from oracle import refmm2
    cores = os.cpu_count() or 1
    P = args.rounds_per_step
    n_pool = min(P, max(args.workers, args.pool))
    pairs = make_pairs(n_pool, 0, args.genome_len)
    sample = args.ref_sample_len
    rounds = [([x[:sample] for x in pairs[j % n_pool][0]], pairs[j % n_pool][1]) for j in range(P)]
    threads = min(cores, 2 * P)
    times, bp = [], 0
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        run_reference_step(refmm2, rounds, threads)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
            bp += sum(len(x) for seqs, _ in rounds for x in seqs)
    total = sum(times)
    value = bp / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8/int32 (SSE2 lanes)", "data": "synthetic",
        "config": {"workload": f"{P} leaf-merge alignment rounds per step, each 2 x {args.genome_len} bp synthetic genomes at 1% "
                               f"divergence, 10 rearrangements (asm10, k=19 w=19)", "sample": f"first {sample} bp of each genome"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"first {sample} bp of each genome of the {P} pairs of a step; the {2 * P} mm_map calls of a step "
                                   f"share {threads} host threads ({cores} cores available)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def free_regs(abi, n_regs, regs):
    tot = 0
    for i in range(len(n_regs)):
        for j in range(n_regs[i]):
            if regs[i][j].p:
                abi._libc.free(C.cast(regs[i][j].p, C.c_void_p))
        if regs[i]:
            abi._libc.free(C.cast(regs[i], C.c_void_p))
        tot += n_regs[i]
    return tot


def pack_regs(n_regs, regs):
    """Match records of one round as bytes (80-byte mm_reg1_t + CIGAR words each), for the gather on rank 0."""
    out = []
    for i in range(len(n_regs)):
        for j in range(n_regs[i]):
            r = regs[i][j]
            out.append(C.string_at(C.addressof(r), 80))
            if r.p:
                out.append(C.string_at(C.addressof(r.p.contents), 24 + 4 * r.p.contents.n_cigar))
    return b"".join(out)


def ours(args):
    import torch
    from pangraph_b200 import abi
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ.setdefault("PGMM_CONTEXTS", str(max(8, args.workers)))
    os.environ.setdefault("PGMM_ARENA_GB", "2")
    L = abi.lib()
    abi.set_device(local)
    from concurrent.futures import ThreadPoolExecutor
    P = args.rounds_per_step
    n_pool = min(P, max(args.workers, args.pool))
    pairs = make_pairs(n_pool, rank * n_pool, args.genome_len)
    bp_pair = [sum(len(x) for x in seqs) for seqs, _ in pairs]
    pool = ThreadPoolExecutor(min(P, args.workers))  # rounds in flight at any moment

    def pairs_of_step(s):
        return [(s * P + j) % n_pool for j in range(P)]

    from pangraph_b200 import sharding

    def round_e2e(p):
        seqs, names = pairs[p]
        idx = abi.Index(seqs, names, "asm10", None, 90)  # mm_idx_str + mm_mapopt_update (host buffers)
        n = len(seqs)
        sa, na = (C.c_char_p * n)(*idx.seqs), (C.c_char_p * n)(*idx.names)
        lens = (C.c_int * n)(*[len(s) for s in idx.seqs])
        n_regs, regs = (C.c_int * n)(), (C.POINTER(abi.mm_reg1_t) * n)()
        L.pgmm_map_batch(idx.mi, n, lens, sa, na, C.byref(idx.mo), n_regs, regs)
        idx.close()
        return n_regs, regs

    def round_resident(idx):
        idx.build()
        out = idx.map_self(raw=True)
        idx.close()  # device blocks go back to the pool for the next round
        return out

    def finish(results):
        """Rank 0 receives the match lists of every rank's rounds, in round order, over NCCL."""
        hits = 0
        if dist is not None:
            payloads = [(rank + world * j, pack_regs(n_regs, regs)) for j, (n_regs, regs) in enumerate(results)]
            got = sharding.gather_rounds(payloads, torch.device("cuda", local))
            assert rank != 0 or len(got) == world * len(results)
        for n_regs, regs in results:
            hits += free_regs(abi, n_regs, regs)
        return hits

    def step_e2e(s):
        return finish(list(pool.map(round_e2e, pairs_of_step(s))))

    def step_resident(idxs):
        return finish(list(pool.map(round_resident, idxs)))

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    cpu_used = {}

    def timed(fn, items):
        sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = os.times()
        t0 = time.perf_counter()
        ev0.record()
        hits = 0
        for it in items:
            hits += fn(it)
        ev1.record()
        sync()
        wall = time.perf_counter() - t0
        c1 = os.times()
        cpu_used[fn.__name__] = ((c1.user - c0.user) + (c1.system - c0.system)) / max(wall, 1e-9)  # busy host cores
        ms = max(ev0.elapsed_time(ev1), 0.0)
        t = torch.tensor([max(wall, ms / 1e3)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), hits

    for s in range(args.warmup):
        step_e2e(s)
    abi.get_stats(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None

    # ---- value: inputs resident in HBM ----
    resident = [[abi.Index(*pairs[p], "asm10", None, 90, resident_only=True) for p in pairs_of_step(args.warmup + s)]
                for s in range(args.steps)]
    abi.get_stats(reset=True)
    t_res, hits_res = timed(step_resident, resident)
    st_res = abi.get_stats(reset=True)
    # ---- e2e: host buffers through the C-ABI ----
    t_e2e, hits_e2e = timed(step_e2e, [args.warmup + s for s in range(args.steps)])
    st_e2e = abi.get_stats(reset=True)
    clocks = sampler.stop() if sampler else None

    bp_rank = sum(bp_pair[p] for s in range(args.steps) for p in pairs_of_step(args.warmup + s))
    bp_all = torch.tensor([bp_rank], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(bp_all)
    bp_total = float(bp_all.item())

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import refmm2
        if os.path.exists(refmm2.REF_SO):
            sample = args.ref_sample_len
            rounds = [([x[:sample] for x in pairs[p][0]], pairs[p][1]) for p in pairs_of_step(0)]
            threads = min(os.cpu_count() or 1, 2 * P)
            t0 = time.perf_counter()
            run_reference_step(refmm2, rounds, threads)
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": sum(len(x) for r in rounds for x in r[0]) / dt / 1e9, "unit": UNIT, "cores": threads,
                            "kind": "reference",
                            "sample": f"one step on the first {sample} bp of each genome ({dt:.1f} s; the {2 * P} mm_map calls share "
                                      f"{threads} host threads, {os.cpu_count()} cores available)"}
    if rank == 0:
        peak, peak_src = measured_peaks()
        # DP kernel families, each timed with CUDA events on the streams its launches go to.  Algorithmic bytes per launch:
        # one traceback byte per in-band cell + the bases each problem reads.  The roofline entry
        # is the family with the largest share of the kernel time (what the ncu launch list shows too).
        fams = {"ksw_extd2_kernel (K5, band-limited extensions)": "k5", "ksw_fill_small_kernel (K5a, first-pass gap fills)": "k5a",
                "ksw_fill_wide_kernel (K5b, long fills across inversions / big indels)": "k5b"}
        dp_kernels = {}
        for name, key in fams.items():
            ms, cells, bases, ln = (st_res[f"{key}_{x}"] for x in ("ms", "cells", "bases", "launches"))
            dp_kernels[name] = {"ms_total": ms, "launches": int(ln), "ms_per_launch": ms / max(1, ln), "cells": cells,
                                "gb_per_s": (cells + bases) / (ms / 1e3) / 1e9 if ms > 0 else 0.0,
                                "gcups": cells / (ms / 1e3) / 1e9 if ms > 0 else 0.0}
        # K4 (chain score fill): one launch per round; algorithmic bytes per anchor = 25 read by the fill (x, y, q_span and
        # the 16-byte window record the prep kernel leaves) + 12 written (f, p, v)
        k4_ms, k4_anchors, k4_launches = st_res["chain_kernel_ms"], st_res["chain_anchors"], st_res["batches"]
        dp_kernels["chain_fill_kernel (K4, chaining score fill)"] = {
            "ms_total": k4_ms, "launches": int(k4_launches), "ms_per_launch": k4_ms / max(1, k4_launches), "anchors": k4_anchors,
            "gb_per_s": 37.0 * k4_anchors / (k4_ms / 1e3) / 1e9 if k4_ms > 0 else 0.0,
            "anchors_per_us": k4_anchors / (k4_ms * 1e3) if k4_ms > 0 else 0.0}
        fams["chain_fill_kernel (K4, chaining score fill)"] = "k4"
        tot_ms = sum(v["ms_total"] for v in dp_kernels.values()) or 1.0
        for v in dp_kernels.values():
            v["share_of_kernel_time"] = v["ms_total"] / tot_ms
        dom = max(dp_kernels, key=lambda k: dp_kernels[k]["ms_total"])
        achieved = dp_kernels[dom]["gb_per_s"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dp_traffic.json")
        if os.path.exists(tp):  # dram bytes per launch from the committed `ncu --set full` captures
            traffic = json.load(open(tp)).get(fams[dom])
        notes = {
            "k4": "one warp per independent anchor segment walks a true dependent chain (each score feeds the next range minimum): "
                  "the launch lasts as long as its longest segment (190 k of a query's 380 k anchors), 212 warp instructions per "
                  "anchor at ~3.4 cycles each, no DRAM traffic to speak of (ncu: 3.9 MB) -- latency-bound, neither roofline applies; "
                  "it is the longest kernel of a round by stream time while using one warp per segment "
                  "(profiles/r01_k4_chain_fill_ncu_full.md)",
            "dp": "the DP kernels are integer-issue bound, not HBM bound (K5a: ~4 instructions per cell against 1 traceback byte; "
                  "ncu: 64 % of issue slots, 4 % of DRAM throughput), and a long fill runs on one SM: the HBM fraction is small "
                  "by construction; launches of concurrent rounds share the GPU"}
        line = {
            "metric": METRIC, "value": bp_total / t_res / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8x4 (DP) / u64 (seeding)", "data": "synthetic",
            "config": {"workload": f"{P} leaf-merge alignment rounds per rank per step, each 2 x {args.genome_len} bp synthetic genomes "
                                   f"at 1% divergence, 10 rearrangements (asm10, k=19 w=19)",
                       "l2": f"working set > L2: {n_pool} distinct genome pairs per rank, rounds in flight work on different pairs, "
                             f"> 1 GB of traceback written per round",
                       "rounds_per_step": P, "rounds_in_flight": min(P, args.workers),
                       "busy_host_cores": {"value": round(cpu_used.get("step_resident", 0), 1), "e2e": round(cpu_used.get("step_e2e", 0), 1)},
                       "hits_per_round": hits_res / max(1, args.steps * P), "host_threads": os.cpu_count()},
            "e2e": {"value": bp_total / t_e2e / 1e9, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "h2d_bytes_per_step": st_e2e["h2d_bytes"] / args.steps, "d2h_bytes_per_step": st_e2e["d2h_bytes"] / args.steps},
            "gpu_launches": int(st_res["launches"]),
            "device_mallocs_in_timed_region": {"value": int(st_res["device_mallocs"]), "e2e": int(st_e2e["device_mallocs"])},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                         "launches": dp_kernels[dom]["launches"], "ms_per_launch": dp_kernels[dom]["ms_per_launch"],
                         "share_of_kernel_time": dp_kernels[dom]["share_of_kernel_time"],
                         "note": notes["k4" if fams[dom] == "k4" else "dp"]},
            # the same figures for the kernel that does most of the GPU's WORK (85 % of all DP cells; CTA trace: 243 of its CTAs
            # resident on average against 14 warps of K4): bound by integer issue, not by HBM
            "roofline_by_gpu_work": {"bound": "hbm", "kernel": "ksw_fill_small_kernel (K5a, first-pass gap fills)",
                                     "achieved": dp_kernels["ksw_fill_small_kernel (K5a, first-pass gap fills)"]["gb_per_s"],
                                     "peak": peak, "unit": "GB/s",
                                     "frac": dp_kernels["ksw_fill_small_kernel (K5a, first-pass gap fills)"]["gb_per_s"] / peak if peak else None,
                                     "traffic": json.load(open(tp)).get("k5a") if os.path.exists(tp) else None,
                                     "note": notes["dp"] + "; launches of concurrent rounds overlap, so the per-launch rate under load "
                                             "is a fraction of the 240 GCUPS a launch reaches alone (profiles/r01_k5a_ncu_full.md)"},
            "kernels": dp_kernels,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "phases_ms_per_round": {k: st_res[k] / (args.steps * P) for k in ("index_ms", "t_encode", "t_seed", "t_chain", "t_dp", "t_stitch",
                                                                       "t_final", "dp_kernel_ms", "total_ms", "t_chain_sort", "t_chain_fill",
                                                                       "t_chain_rest", "chain_kernel_ms")},
            "chain": {"anchors_per_round": st_res["chain_anchors"] / (args.steps * P), "segments_per_round": st_res["chain_segments"] / (args.steps * P),
                      "segments_to_host_arbiter": st_res["chain_redo_segments"] / (args.steps * P),
                      "anchors_to_host_arbiter": st_res["chain_redo_anchors"] / (args.steps * P),
                      "fill_kernel_ms_per_round": st_res["chain_kernel_ms"] / (args.steps * P)},
            "dp": {"jobs_per_round": st_res["dp_jobs"] / (args.steps * P), "cells_per_round": st_res["dp_cells"] / (args.steps * P),
                   "waves_per_round": st_res["dp_waves"] / (args.steps * P)},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--rounds-per-step", type=int, default=192, help="independent leaf-merge rounds a rank keeps in flight per step")
    ap.add_argument("--workers", type=int, default=64, help="host threads driving rounds concurrently (one CUDA stream each)")
    ap.add_argument("--pool", type=int, default=36, help="distinct genome pairs generated per rank (steps cycle through them)")
    ap.add_argument("--ref-sample-len", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
