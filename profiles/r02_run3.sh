set -x
python profiles/k6_probe.py 80 125000 16 > gpurun_out/k6_probe.txt 2>&1; cat gpurun_out/k6_probe.txt
python profiles/k6_probe.py 2000 20000 16 > gpurun_out/k6_probe_many.txt 2>&1; cat gpurun_out/k6_probe_many.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nextalign -c 2 -f -o gpurun_out/r02_k6 python profiles/k6_probe.py 2000 20000 2 > gpurun_out/ncu_k6.log 2>&1
ncu -i gpurun_out/r02_k6.ncu-rep --page raw --csv > gpurun_out/r02_k6_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
