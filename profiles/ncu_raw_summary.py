"""`ncu -i x.ncu-rep --page raw --csv` -> the metrics this repo quotes, as a markdown table per launch.  usage: ncu_raw_summary.py raw.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "lts__t_sector_hit_rate.pct"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
for d in data:
    g = lambda k: d[hdr.index(k)]
    print(f"## `{g('Kernel Name')[:60]}` grid {g('Grid Size')} block {g('Block Size')}\n\n| metric | value | unit |\n|---|---:|---|")
    for w in WANT:
        if w in hdr:
            print(f"| {w} | {g(w)} | {units[hdr.index(w)]} |")
    print()
