#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02_d.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_d.txt
timeout 900 python bench.py > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err
echo "bench rc=$?" >> gpurun_out/bench_r02_d.err
tail -12 gpurun_out/gpu_tests_r02_d.txt; tail -3 gpurun_out/bench_r02_d.err
