#!/bin/bash
# first GPU pass of round 2: parity suite, kernel-variant timings, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_a_smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02_a.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_a.txt
for v in default byval; do
  if [ "$v" = default ]; then unset IBFT_LIB; else export IBFT_LIB=$PWD/go-ibft_b200/variants/lib_$v.so; fi
  timeout 300 python tools/quick_bench.py 20 >> gpurun_out/quick_r02_a.jsonl 2>> gpurun_out/quick_r02_a.err
done
unset IBFT_LIB
timeout 600 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err
tail -3 gpurun_out/gpu_tests_r02_a.txt; cat gpurun_out/quick_r02_a.jsonl; tail -c 600 gpurun_out/bench_r02_a.json
