#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02_e.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_e.txt
rm -f gpurun_out/quick_r02_variants.jsonl
for v in default blind1 blind2 byval default; do
  if [ "$v" = default ]; then unset IBFT_LIB; else export IBFT_LIB=$PWD/go-ibft_b200/variants/lib_$v.so; fi
  timeout 300 python tools/quick_bench.py 20 >> gpurun_out/quick_r02_variants.jsonl 2>> gpurun_out/quick_r02_variants.err
done
unset IBFT_LIB
timeout 900 python bench.py > gpurun_out/bench_r02_e.json 2> gpurun_out/bench_r02_e.err
echo "bench rc=$?" >> gpurun_out/bench_r02_e.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_keycache.py::test_known_key_latency_path_on_a_10k_round" "tests/test_gpu_round2.py::test_two_span_payload_kind_matches_single_span" -x -q > gpurun_out/sanitizer_memcheck_r02_known_latency.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_r02_known_latency.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_keycache.py::test_known_key_latency_path_on_a_10k_round" -x -q > gpurun_out/sanitizer_racecheck_r02_known_latency.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_r02_known_latency.txt
tail -8 gpurun_out/gpu_tests_r02_e.txt; cat gpurun_out/quick_r02_variants.jsonl; tail -2 gpurun_out/bench_r02_e.err; tail -4 gpurun_out/sanitizer_memcheck_r02_known_latency.txt; tail -4 gpurun_out/sanitizer_racecheck_r02_known_latency.txt
