"""Throughput sweeps over the benchmark's knobs; each line: label, value, e2e, ms/step, busy host cores, ms per round (total, chain fill, dp).
usage: sweep.py "label:ENV=V,ENV2=V:--flag v --flag2 v" ..."""
import json
import os
import subprocess
import sys

for spec in sys.argv[1:]:
    label, envs, flags = (spec.split(":") + ["", ""])[:3]
    env = dict(os.environ)
    for kv in filter(None, envs.split(",")):
        k, v = kv.split("=")
        env[k] = v
    r = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline"] + flags.split(), env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        ph = d["phases_ms_per_round"]
        print(label, round(d["value"], 4), round(d["e2e"]["value"], 4), round(d["ms_per_step"]), d["config"]["busy_host_cores"],
              {k: round(ph[k]) for k in ("total_ms", "t_seed", "t_chain_fill", "t_chain_rest", "t_dp", "t_stitch")}, flush=True)
    except Exception as e:  # noqa: BLE001
        print(label, "FAILED", e, r.stderr[-400:], flush=True)
