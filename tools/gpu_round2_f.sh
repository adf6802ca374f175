#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02_f.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_f.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_f.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_f.txt
timeout 900 python bench.py > gpurun_out/bench_r02_f.json 2> gpurun_out/bench_r02_f.err
echo "bench rc=$?" >> gpurun_out/bench_r02_f.err
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_keycache.py::test_known_key_latency_path_on_a_10k_round" "tests/test_gpu_verify.py::test_ragged_sizes" -x -q > gpurun_out/sanitizer_racecheck_r02_known_latency.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_r02_known_latency.txt
tail -6 gpurun_out/gpu_tests_r02_f.txt; tail -2 gpurun_out/smoke_r02_f.txt; tail -2 gpurun_out/bench_r02_f.err; tail -4 gpurun_out/sanitizer_racecheck_r02_known_latency.txt
