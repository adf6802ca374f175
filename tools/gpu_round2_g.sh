#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02_g.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_g.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_g.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_g.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_g_reference.json 2> gpurun_out/bench_r02_g_reference.err
timeout 900 python bench.py > gpurun_out/bench_r02_g.json 2> gpurun_out/bench_r02_g.err
echo "bench rc=$?" >> gpurun_out/bench_r02_g.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02_v3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-key-cache-leg --latency-reps 3 > gpurun_out/bench_under_ncu_r02_v3.log 2>&1
tail -5 gpurun_out/gpu_tests_r02_g.txt; tail -2 gpurun_out/smoke_r02_g.txt; tail -2 gpurun_out/bench_r02_g.err; wc -l gpurun_out/launches_r02_v3.csv
