#!/bin/bash
# round-2 verification run on the comb-registry / peer-exchange build: full GPU suite, smoke, sanitizer on the new kernels,
# ncu capture of k_verify_known, launch list of the bench, both bench arms
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02_v4.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_v4.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_v4.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_v4.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_keycache.py -x -q -k "carry_over or (learn_then_verify and config2)" > gpurun_out/sanitizer_memcheck_r02_v4_keycache.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_keycache.py -x -q -k "carry_over" > gpurun_out/sanitizer_racecheck_r02_v4_keycache.txt 2>&1
KEY_CACHE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_verify_known -s 3 -c 1 -f -o gpurun_out/ncu_known_r02_v1 \
  python tools/quick_bench.py 20 > gpurun_out/ncu_known_r02_v1.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_v6_reference_n1.json 2> gpurun_out/bench_r02_v6_reference_n1.err
timeout 900 python bench.py > gpurun_out/bench_r02_v6_n1.json 2> gpurun_out/bench_r02_v6_n1.err
echo "bench rc=$?" >> gpurun_out/bench_r02_v6_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02_v4.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latency-reps 3 > gpurun_out/bench_under_ncu_r02_v4.log 2>&1
tail -5 gpurun_out/gpu_tests_r02_v4.txt; tail -2 gpurun_out/smoke_r02_v4.txt; tail -2 gpurun_out/bench_r02_v6_n1.err
grep -h "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_*_r02_v4_keycache.txt | head
ls -la gpurun_out/ncu_known_r02_v1.ncu-rep; wc -l gpurun_out/launches_r02_v4.csv
