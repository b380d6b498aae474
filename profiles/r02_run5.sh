set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gputest5.log 2>&1; tail -4 gpurun_out/gputest5.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sketch_tile -c 2 -f -o gpurun_out/r02_k1 python bench.py --steps 1 --warmup 0 --rounds-per-step 1 --workers 1 --pool 1 --no-parity --no-cpu-baseline --solo-rounds 0 > gpurun_out/ncu_k1.log 2>&1
ncu -i gpurun_out/r02_k1.ncu-rep --page raw --csv > gpurun_out/r02_k1_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_k1.csv python bench.py --steps 1 --warmup 0 --rounds-per-step 1 --workers 1 --pool 1 --no-parity --no-cpu-baseline --solo-rounds 0 > gpurun_out/ncu_k1b.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_k1.log 2> gpurun_out/bench_k1.err; tail -2 gpurun_out/bench_k1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_k1.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["busy_host_cores"])
print(d["phases_ms_per_round"]); print(d["host_cpu_ms_per_round"])
PY
