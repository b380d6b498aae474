set -x
python profiles/cta_trace_run.py 64 192 > gpurun_out/cta_trace_r02.txt 2>&1
tail -12 gpurun_out/cta_trace_r02.txt
python profiles/sweep.py "w32::--steps 3 --warmup 3 --workers 32" "w96::--steps 3 --warmup 3 --workers 96" "w128c32::--steps 3 --warmup 3 --workers 128 --contexts 32" > gpurun_out/sweep_r02.txt 2>&1
cat gpurun_out/sweep_r02.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 0 --rounds-per-step 2 --workers 2 --pool 2 --no-parity --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
wc -l gpurun_out/r02_launches.csv
rm -f gpurun_out/cta_trace.npy
