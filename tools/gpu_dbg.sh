#!/bin/bash
# comb key registry + carry-over: tests, memcheck of the new kernels, full default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_keycache.py tests/test_gpu_host.py -x -q > gpurun_out/gpu_tests_r02_comb.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r02_comb.txt
tail -15 gpurun_out/gpu_tests_r02_comb.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_keycache.py -x -q -k "carry_over or learn_then_verify and config2" > gpurun_out/sanitizer_memcheck_r02_comb.txt 2>&1
grep -n "ERROR SUMMARY\|passed\|failed\|Invalid" gpurun_out/sanitizer_memcheck_r02_comb.txt | head
timeout 900 python bench.py > gpurun_out/bench_r02_v5_n1.json 2> gpurun_out/bench_r02_v5_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_v5_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
print(json.dumps(d.get('known_validator_path'), indent=1)[:2500])
PY
