#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_configs.py::test_sharded_verifier_pipeline_single_rank" -x -q > gpurun_out/dbg_p2p.txt 2>&1
echo "plain rc=$?" >> gpurun_out/dbg_p2p.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_configs.py::test_sharded_verifier_pipeline_single_rank" -x -q > gpurun_out/dbg_p2p_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/dbg_p2p_memcheck.txt
tail -15 gpurun_out/dbg_p2p.txt; grep -n "Invalid\|at \|by \|=========" gpurun_out/dbg_p2p_memcheck.txt | head -30
