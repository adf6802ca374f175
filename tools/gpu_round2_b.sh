#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02_b.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_b.txt
timeout 900 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err
echo "bench rc=$?" >> gpurun_out/bench_r02_b.err
tail -15 gpurun_out/gpu_tests_r02_b.txt; tail -5 gpurun_out/bench_r02_b.err; tail -c 3000 gpurun_out/bench_r02_b.json
