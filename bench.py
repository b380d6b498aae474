#!/usr/bin/env python
"""Benchmark of the alignment hot path of `pangraph build` (BASELINE.json metric: Gbp aligned / s).

One ROUND = one find_matches call of a guide-tree leaf merge: two related synthetic 5-Mbp genomes (1 % divergence,
10 rearrangements each) are indexed and aligned all-vs-all, i.e. mm_idx_str + mm_mapopt_update + one
mm_map per sequence in the reference, index kernels + pgmm_map_batch here.
One STEP = `--rounds-per-step` (default 192) such rounds, `--workers` of them in flight at any moment (default: 6 per host core of the rank, 12..64): sibling leaf merges of the guide tree are independent
(merge_graphs only reads its two children), so a rank keeps several of them in flight, one host thread and one CUDA
stream each; their DP waves are merged across rounds by the library's DP service (dp_service.cu).  bp per step = total length of the genomes of its rounds.

  python bench.py [--gpus N --steps K --warmup W]        our CUDA path (N>1: one rank per GPU under torchrun; every
                                                          rank aligns its own pairs -- leaf merges are independent --
                                                          and the match lists are gathered on rank 0 over NCCL)
  python bench.py --impl reference [...]                  the reference's own C (oracle/_ref) on the host cores

`value`  : inputs resident in HBM before the timed region (one pgmm_idx_upload per round of a step, done ONCE and reused by
           every step -- device memory is O(rounds per step), independent of --steps; timed: index kernels + pgmm_map_self)
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
    from concurrent.futures import ThreadPoolExecutor
    from pangraph_b200 import synth
    anc = synth.ancestor(length, 42)
    ids = [2 * p + h for p in range(first_pair, first_pair + n_pairs) for h in (0, 1)]
    with ThreadPoolExecutor(4) as ex:  # numpy drops the GIL in the large array operations
        gs = list(ex.map(lambda i: synth.mutate(anc, 20260 + i).tobytes(), ids))
    return [([gs[2 * j], gs[2 * j + 1]], [str(ids[2 * j]), str(ids[2 * j + 1])]) for j in range(n_pairs)]


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


def reference_round(refmm2, lib, seqs, names):
    """One find_matches round the way the reference runs it: mm_idx_str + mm_mapopt_update, then mm_map per sequence."""
    ix = refmm2.Index(lib, seqs, names, "asm10", None, 90)
    try:
        return [ix.map_one(i) for i in range(len(ix.seqs))]
    finally:
        ix.close()


def run_reference_step(refmm2, rounds, threads):
    """The reference's path for every round of a step; rounds share one pool of `threads` host threads (sibling merges
    of the guide tree run concurrently in the reference too; ctypes drops the GIL inside the C calls)."""
    from concurrent.futures import ThreadPoolExecutor
    lib = refmm2.load_ref()
    with ThreadPoolExecutor(max(1, threads)) as ex:
        return list(ex.map(lambda r: sum(len(h) for h in reference_round(refmm2, lib, r[0], r[1])), rounds))


def reference_sample(args, pairs, cores):
    """The bounded sample of the workload the CPU arm runs per step: FULL-SIZE rounds of the same pair family (same bytes
    per round as the CUDA arm), only fewer of them: 2 per host thread."""
    n = args.ref_rounds_per_step if args.ref_rounds_per_step > 0 else 2 * cores
    n = max(1, min(n, args.rounds_per_step))
    return [pairs[j % len(pairs)] for j in range(n)]


def reference_arm(args):
    """The reference's own CPU implementation (oracle/_ref = its vendored minimap2 C, unmodified) on a bounded sample of
    the same workload, with all host threads it can use."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from oracle import refmm2
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    P = args.rounds_per_step
    n_ref = args.ref_rounds_per_step if args.ref_rounds_per_step > 0 else 2 * cores
    pairs = make_pairs(min(args.pool, max(1, min(n_ref, P))), 0, args.genome_len)
    rounds = reference_sample(args, pairs, cores)
    threads = min(cores, len(rounds))
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
    sample = (f"{len(rounds)} full-size rounds per step (2 x {args.genome_len} bp each, the same pair family and the same bytes per "
              f"round as the CUDA arm's {P} rounds per step), one host thread per round in flight, {threads} threads "
              f"({cores} cores available)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8/int32 (SSE2 lanes)", "data": "synthetic",
        "config": {"workload": f"{P} leaf-merge alignment rounds per rank per step, each 2 x {args.genome_len} bp synthetic genomes at 1% "
                               f"divergence, 10 rearrangements (asm10, k=19 w=19)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
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


def regs_to_tuples(abi, n_regs, regs):
    """Every field of every hit of one round, like oracle.refmm2.reg_to_tuple (parity gate only; does not free)."""
    return [[abi.reg_to_tuple(regs[i][j]) for j in range(n_regs[i])] for i in range(len(n_regs))]


def device_footprint_gb(rounds_per_step, workers, genome_len, steps=None, dp_arena_frac=0.4, total_gb=178.0):
    """Upper estimate of the device memory one rank of the CUDA arm holds (GB).  Independent of --steps by construction:
    the resident arm keeps one uploaded pair per round OF A STEP and reuses it every step.
      resident pairs: rounds_per_step x (coded bases 1.25 B/bp + index and minimizer arrays ~5 B/bp)
      contexts      : workers x ~110 B/bp of seeding / chaining workspace for one round (2 genomes)
      DP service    : at most dp_arena_frac of what is free when it starts (arenas grow on demand)."""
    bp = 2.0 * genome_len
    resident = rounds_per_step * bp * (1.25 + 5.0) / 1e9
    contexts = workers * bp * 110.0 / 1e9
    arenas = dp_arena_frac * max(0.0, total_gb - resident * 0.2 - contexts * 0.5)
    return {"resident_gb": resident, "contexts_gb": contexts, "dp_arenas_gb": arenas, "total_gb": resident + contexts + arenas,
            "depends_on_steps": False}


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
    # execution contexts (seeding / chaining workspaces, ~1 GB each) are held for the first third of a round only
    os.environ.setdefault("PGMM_CONTEXTS", str(max(8, min(args.workers, args.contexts))))
    os.environ.setdefault("PGMM_ARENA_GB", "2")
    L = abi.lib()
    abi.set_device(local)
    from concurrent.futures import ThreadPoolExecutor
    P = args.rounds_per_step
    n_pool = max(1, min(P, args.pool))
    pairs = make_pairs(n_pool, rank * n_pool, args.genome_len)
    bp_pair = [sum(len(x) for x in seqs) for seqs, _ in pairs]
    pool = ThreadPoolExecutor(min(P, args.workers))  # rounds in flight at any moment

    def pairs_of_step(s):
        return [(s * P + j) % n_pool for j in range(P)]

    from pangraph_b200 import sharding

    def round_e2e(p, keep=False):
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
        idx.build()  # K1 + K2 on the resident bases + mm_mapopt_update; rebuilds in place every step
        return idx.map_self(raw=True)

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

    # The timed region feeds the rounds of ALL its steps to the pool at once: a step's results are collected (and
    # gathered on rank 0) as soon as its own rounds are done, while rounds of the following steps are already running --
    # the pipeline is not drained between steps, only at the two ends of the timed region.
    step_e2e.submit = lambda s: [pool.submit(round_e2e, p) for p in pairs_of_step(s)]
    res_locks = {}

    def round_resident_locked(idx):
        with res_locks.setdefault(id(idx), threading.Lock()):  # a resident pair is rebuilt in place: one round at a time
            return round_resident(idx)

    step_resident.submit = lambda idxs: [pool.submit(round_resident_locked, ix) for ix in idxs]

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- parity gate (outside every timed region): round 0 of this rank through BOTH call paths against the reference's
    # own C (oracle/_ref).  A run whose results differ does not produce a number.
    parity = {"parity_checked": False}
    from oracle import refmm2
    have_ref = os.path.exists(refmm2.REF_SO)
    if have_ref and not args.no_parity and (rank == 0 or args.parity_all_ranks):
        seqs, names = pairs[0]
        t0 = time.perf_counter()
        want, mid = refmm2.ref_map_all(seqs, names, "asm10", None, 90, threads=2)
        t_ref = time.perf_counter() - t0
        n_regs, regs = round_e2e(0)
        got_e2e = regs_to_tuples(abi, n_regs, regs)
        free_regs(abi, n_regs, regs)
        ix = abi.Index(*pairs[0], "asm10", None, 90, resident_only=True)
        n_regs, regs = round_resident(ix)
        got_res = regs_to_tuples(abi, n_regs, regs)
        free_regs(abi, n_regs, regs)
        assert ix.mo.mid_occ == mid, f"mid_occ {ix.mo.mid_occ} != reference {mid}"
        ix.close()
        if got_e2e != want or got_res != want:
            raise SystemExit("bench.py: PARITY FAILURE -- the CUDA path's hits differ from the reference's on round 0; no number reported")
        parity = {"parity_checked": True, "parity_round": f"pair 0 of rank {rank} (2 x {args.genome_len} bp), host-buffer and resident call "
                  f"paths, every mm_reg1_t field + CIGAR + de equal to oracle/_ref ({sum(len(w) for w in want)} hits)",
                  "reference_seconds_one_round_2_threads": round(t_ref, 2)}
    elif not have_ref:
        parity["parity_note"] = "oracle/_ref/libmm2ref.so not present on this box"

    cpu_used = {}

    def timed(fn, items):
        sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = os.times()
        t0 = time.perf_counter()
        ev0.record()
        hits = 0
        futures = [fn.submit(it) for it in items]
        for fs in futures:
            hits += finish([f.result() for f in fs])
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

    # ---- inputs of the resident arm: ONE uploaded pair per round of a step, reused by every step (O(P) device memory) ----
    resident = [abi.Index(*pairs[p], "asm10", None, 90, resident_only=True) for p in pairs_of_step(0)]
    for s in range(args.warmup):
        if s % 2 == 0:
            step_e2e(s)
        else:
            step_resident(resident)
    abi.get_stats(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None

    # ---- value: inputs resident in HBM ----
    free0 = torch.cuda.mem_get_info()[0]
    t_res, hits_res = timed(step_resident, [resident] * args.steps)
    st_res = abi.get_stats(reset=True)
    # ---- e2e: host buffers through the C-ABI ----
    t_e2e, hits_e2e = timed(step_e2e, [args.warmup + s for s in range(args.steps)])
    st_e2e = abi.get_stats(reset=True)
    clocks = sampler.stop() if sampler else None
    free1, total_mem = torch.cuda.mem_get_info()

    # ---- solo pass (not part of any timed region): a few rounds one at a time, nothing else on the GPU, so that the
    # CUDA-event duration of a launch is the kernel's own time (under load the launches of up to 64 rounds share the SMs
    # and wait in hardware queues: their event times say how long a wave takes, not how fast the kernel is) ----
    sync()
    for j in range(args.solo_rounds):
        n_regs, regs = round_resident(resident[j % len(resident)])
        free_regs(abi, n_regs, regs)
    sync()
    st_solo = abi.get_stats(reset=True)

    bp_rank = args.steps * sum(bp_pair[p] for p in pairs_of_step(0))
    bp_rank_e2e = sum(bp_pair[p] for s in range(args.steps) for p in pairs_of_step(args.warmup + s))
    bp_all = torch.tensor([bp_rank, bp_rank_e2e], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(bp_all)
    bp_total, bp_total_e2e = float(bp_all[0].item()), float(bp_all[1].item())

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and have_ref:
        cores = os.cpu_count() or 1
        try:
            cores = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            pass
        rounds = reference_sample(args, pairs, cores)
        threads = min(cores, len(rounds))
        t0 = time.perf_counter()
        run_reference_step(refmm2, rounds, threads)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": sum(len(x) for r in rounds for x in r[0]) / dt / 1e9, "unit": UNIT, "cores": threads,
                        "kind": "reference",
                        "sample": f"{len(rounds)} full-size rounds of the same pair family (2 x {args.genome_len} bp each, same bytes per "
                                  f"round as the CUDA arm), one host thread per round in flight, {threads} threads, {dt:.1f} s "
                                  f"({cores} cores available)"}
    if rank == 0:
        peak, peak_src = measured_peaks()
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        int_peak = n_sm * 128 * sm_mhz * 1e6 / 1e9  # G int32-lane-ops/s at the clock sampled under load
        # DP kernel families, each launch timed with CUDA events on the stream it is launched on.  Algorithmic work per
        # launch: 50 integer lane-ops per cell and, for the HBM reading, one traceback byte per in-band cell
        # plus the bases each problem reads.
        fams = {"ksw_extd2_kernel (K5, band-limited extensions)": "k5", "ksw_fill_small_kernel (K5a, first-pass gap fills)": "k5a",
                "ksw_fill_wide_kernel (K5b, long fills across inversions / big indels)": "k5b"}
        tp = os.path.join(ROOT, "profiles", "dp_traffic.json")
        traffic_tbl = json.load(open(tp)) if os.path.exists(tp) else {}
        def family(st, key):
            ms, cells, bases, ln = (st[f"{key}_{x}"] for x in ("ms", "cells", "bases", "launches"))
            sec = ms / 1e3
            return {"ms_total": ms, "launches": int(ln), "ms_per_launch": ms / max(1, ln), "cells": cells,
                    "cells_per_launch": cells / max(1, ln),
                    "gcups": cells / sec / 1e9 if ms > 0 else 0.0,
                    "achieved_int_gops": 50.0 * cells / sec / 1e9 if ms > 0 else 0.0,
                    "int_frac": 50.0 * cells / sec / 1e9 / int_peak if ms > 0 else 0.0,
                    "gb_per_s": (cells + bases) / sec / 1e9 if ms > 0 else 0.0,
                    "hbm_frac": (cells + bases) / sec / 1e9 / peak if ms > 0 else 0.0}

        dp_kernels = {}
        for name, key in fams.items():
            dp_kernels[name] = family(st_res, key)  # under load: launches of all rounds in flight share the GPU
            dp_kernels[name]["traffic_bytes_per_launch_ncu"] = traffic_tbl.get(key)
            dp_kernels[name]["solo"] = family(st_solo, key)  # one round at a time: the kernel's own duration
        # K4 (chain score fill): latency-bound dependent chain; reported as anchors/us.  Algorithmic bytes per anchor = 25
        # read by the fill (x, y, q_span and the 16-byte window record of the prep kernel) + 12 written (f, p, v)
        k4_ms, k4_anchors, k4_launches = st_res["chain_kernel_ms"], st_res["chain_anchors"], max(1.0, st_res["chain_launches"])
        k4 = {"ms_total": k4_ms, "launches": int(k4_launches), "ms_per_launch": k4_ms / k4_launches, "anchors": k4_anchors,
              "bound": "latency (K4p: ~20 short launches per fixed-point iteration, pointer jumping over the predecessor forest)",
              "anchors_per_us": k4_anchors / (k4_ms * 1e3) if k4_ms > 0 else 0.0,
              "gb_per_s": 37.0 * k4_anchors / (k4_ms / 1e3) / 1e9 if k4_ms > 0 else 0.0,
              "solo": {"ms_per_round": st_solo["chain_kernel_ms"] / max(1, args.solo_rounds), "launches_per_round": st_solo["chain_launches"] / max(1, args.solo_rounds),
                       "anchors_per_us": st_solo["chain_anchors"] / (st_solo["chain_kernel_ms"] * 1e3) if st_solo["chain_kernel_ms"] > 0 else 0.0}}
        tot_ms = (sum(v["ms_total"] for v in dp_kernels.values()) + k4_ms) or 1.0
        tot_cells = sum(v["cells"] for v in dp_kernels.values()) or 1.0
        for v in list(dp_kernels.values()) + [k4]:
            v["share_of_kernel_time"] = v["ms_total"] / tot_ms
        for v in dp_kernels.values():
            v["share_of_cells"] = v["cells"] / tot_cells
        dom = max(dp_kernels, key=lambda k: dp_kernels[k]["cells"])  # the kernel that does most of the DP cells
        d_load = dp_kernels[dom]
        d = dict(d_load["solo"]) if d_load["solo"]["ms_total"] > 0 else dict(d_load)
        d["share_of_cells"], d["share_of_kernel_time"] = d_load["share_of_cells"], d_load["share_of_kernel_time"]
        d["traffic_bytes_per_launch_ncu"] = d_load["traffic_bytes_per_launch_ncu"]
        rounds_timed = args.steps * P
        line = {
            "metric": METRIC, "value": bp_total / t_res / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16x2 / int8x4 lanes (DP), u64 (seeding), f32+f64 (chain scores)", "data": "synthetic",
            "config": {"workload": f"{P} leaf-merge alignment rounds per rank per step, each 2 x {args.genome_len} bp synthetic genomes "
                                   f"at 1% divergence, 10 rearrangements (asm10, k=19 w=19)",
                       "l2": f"working set > L2: {n_pool} distinct genome pairs per rank ({n_pool * 2 * args.genome_len / 1e6:.0f} MB of bases), "
                             f"rounds in flight work on different pairs, > 1 GB of traceback written per round",
                       "rounds_per_step": P, "rounds_in_flight": min(P, args.workers), "host_cores_per_rank": (os.cpu_count() or 1) // max(1, env_int("LOCAL_WORLD_SIZE", world)), "distinct_pairs_per_rank": n_pool,
                       "steps_pipelined": "rounds of consecutive steps overlap inside the timed region (no drain between steps)",
                       "busy_host_cores": {"value": round(cpu_used.get("step_resident", 0), 1), "e2e": round(cpu_used.get("step_e2e", 0), 1)},
                       "hits_per_round": hits_res / max(1, rounds_timed), "host_threads": os.cpu_count(),
                       "device_memory_gb": {"total": total_mem / 1e9, "free_before_timed": free0 / 1e9, "free_after_timed": free1 / 1e9},
                       "device_footprint_model": device_footprint_gb(P, min(P, args.workers, args.contexts), args.genome_len)},
            "e2e": {"value": bp_total_e2e / t_e2e / 1e9, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "h2d_bytes_per_step": st_e2e["h2d_bytes"] / args.steps, "d2h_bytes_per_step": st_e2e["d2h_bytes"] / args.steps},
            "gpu_launches": int(st_res["launches"]),
            "device_mallocs_in_timed_region": {"value": int(st_res["device_mallocs"]), "e2e": int(st_e2e["device_mallocs"])},
            # headline roofline: the DP kernel that does most cells.  The DP is integer-issue bound (ncu: K5a 64 % of issue
            # slots, 4 % of DRAM throughput), so `achieved`/`peak` are integer lane-ops; the HBM reading north_star asks
            # for is given next to it and is small by construction (1 traceback byte per ~50 lane-ops).
            "roofline": {"bound": "int-issue", "kernel": dom, "achieved": d["achieved_int_gops"], "peak": int_peak,
                         "unit": "G int32-lane-ops/s", "frac": d["int_frac"],
                         "definition": "achieved = 50 lane-ops x cells of the family / its summed per-launch CUDA-event time, launches timed "
                                       f"with CUDA events on their own stream in a solo pass of {args.solo_rounds} rounds after the timed regions (one round at "
                                       f"a time, so a launch has the GPU to itself like in the ncu launch list); peak = {n_sm} SMs x 128 lanes x "
                                       f"{sm_mhz:.0f} MHz (median SM clock sampled during the timed region)",
                         "gcups": d["gcups"], "launches": d["launches"], "ms_per_launch": d["ms_per_launch"],
                         "cells_per_launch": d["cells_per_launch"], "share_of_cells": d["share_of_cells"],
                         "share_of_kernel_time": d["share_of_kernel_time"],
                         "hbm": {"achieved": d["gb_per_s"], "peak": peak, "unit": "GB/s", "frac": d["hbm_frac"], "peak_source": peak_src,
                                 "algorithmic_bytes": "1 traceback byte per in-band cell + the bases each problem reads"},
                         "traffic": d["traffic_bytes_per_launch_ncu"],
                         "under_load": {"gcups": d_load["gcups"], "ms_per_launch": d_load["ms_per_launch"], "launches": d_load["launches"],
                                        "cells_per_launch": d_load["cells_per_launch"], "int_frac": d_load["int_frac"],
                                        "note": "event times of launches inside the timed region: up to 64 rounds' launches share the SMs and "
                                                "wait in hardware queues, so this measures wave latency, not kernel speed"}},
            "kernels": dp_kernels,
            "chain_fill_kernel (K4)": k4,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "phases_ms_per_round": {k: st_res[k] / rounds_timed for k in ("index_ms", "t_encode", "t_seed", "t_chain", "t_dp", "t_stitch",
                                                                       "t_final", "dp_kernel_ms", "total_ms", "t_chain_sort", "t_chain_fill",
                                                                       "t_chain_rest", "chain_kernel_ms")},
            "host_cpu_ms_per_round": {k[4:]: st_res[k] / rounds_timed for k in st_res if k.startswith("cpu_")},
            "chain": {"anchors_per_round": st_res["chain_anchors"] / rounds_timed, "segments_per_round": st_res["chain_segments"] / rounds_timed,
                      "segments_to_host_arbiter": st_res["chain_redo_segments"] / rounds_timed,
                      "anchors_to_host_arbiter": st_res["chain_redo_anchors"] / rounds_timed,
                      "fill_kernel_ms_per_round": st_res["chain_kernel_ms"] / rounds_timed},
            "dp": {"jobs_per_round": st_res["dp_jobs"] / rounds_timed, "cells_per_round": st_res["dp_cells"] / rounds_timed,
                   "cells_per_bp": st_res["dp_cells"] / max(1.0, bp_rank), "waves_per_round": st_res["dp_waves"] / rounds_timed},
        }
        line.update(parity)
        print(json.dumps(line))
    for ix in resident:
        ix.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--rounds-per-step", type=int, default=192, help="independent leaf-merge rounds of one rank per step")
    ap.add_argument("--workers", type=int, default=0, help="rounds in flight per rank (one host thread each); 0 = 6 per host core of the rank, at least 12, at most 64")
    ap.add_argument("--contexts", type=int, default=24, help="execution contexts of the library (PGMM_CONTEXTS)")
    ap.add_argument("--pool", type=int, default=32, help="distinct genome pairs generated per rank (rounds cycle through them)")
    ap.add_argument("--ref-rounds-per-step", type=int, default=0, help="reference arm / cpu_baseline: full-size rounds per step (0 = 2 per host thread)")
    ap.add_argument("--solo-rounds", type=int, default=4, help="rounds run one at a time after the timed regions to time each kernel family alone")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate against oracle/_ref (profiling runs only)")
    ap.add_argument("--parity-all-ranks", action="store_true")
    args = ap.parse_args()
    if args.workers <= 0:
        # With few host cores per GPU (the 8-GPU boxes of this pool have 4) the host is the bound and more rounds in flight only
        # add contention: 4 cores, 24 in flight 0.505 Gbp/s end to end against 0.421 with 64 (taskset on a one-GPU box).
        try:
            cores = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            cores = os.cpu_count() or 1
        per_rank = max(1, cores // max(1, env_int("LOCAL_WORLD_SIZE", env_int("WORLD_SIZE", 1))))
        args.workers = max(12, min(64, 6 * per_rank))
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
