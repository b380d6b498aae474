set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gputest4.log 2>&1; tail -4 gpurun_out/gputest4.log
python profiles/sweep.py "base::--steps 4 --warmup 3 --no-parity" "nt3:PGMM_MAX_NT_TIER=3:--steps 4 --warmup 3 --no-parity" "nt2:PGMM_MAX_NT_TIER=2:--steps 4 --warmup 3 --no-parity" > gpurun_out/sweep_nt.txt 2>&1
cat gpurun_out/sweep_nt.txt
