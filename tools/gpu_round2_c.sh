#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02_c.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_c.txt
timeout 900 python bench.py > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err
echo "bench rc=$?" >> gpurun_out/bench_r02_c.err
# launch list of the same command (shares, not absolutes) and one full capture of the dominant kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_v1.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-key-cache-leg --latency-reps 3 > gpurun_out/bench_under_ncu_r02.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_recover -s 3 -c 1 -f -o gpurun_out/ncu_recover_r02_v1 \
  python tools/quick_bench.py 20 > gpurun_out/ncu_quick_r02.log 2>&1
tail -12 gpurun_out/gpu_tests_r02_c.txt; tail -3 gpurun_out/bench_r02_c.err; ls -la gpurun_out | tail -5
