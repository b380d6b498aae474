set -x
python -m pytest tests/test_gpu_ksw.py tests/test_gpu_map.py tests/test_golden.py -m gpu -x -q > gpurun_out/gputest3.log 2>&1; tail -3 gpurun_out/gputest3.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_walk.log 2> gpurun_out/bench_walk.err; tail -3 gpurun_out/bench_walk.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_walk.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["busy_host_cores"])
for k,v in d["kernels"].items(): print(k, round(v["ms_total"]), v["launches"], round(v["ms_per_launch"],2), round(v["gcups"],2))
print(d["phases_ms_per_round"])
PY
for K in ksw_fill_small par_decide; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -c 2 -f -o gpurun_out/r02_$K python bench.py --steps 1 --warmup 0 --rounds-per-step 1 --workers 1 --pool 1 --no-parity --no-cpu-baseline > gpurun_out/ncu_$K.log 2>&1
ncu -i gpurun_out/r02_$K.ncu-rep --page raw --csv > gpurun_out/r02_${K}_raw.csv 2>/dev/null
done
ls -la gpurun_out/
