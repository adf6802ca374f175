#!/bin/bash
# occupancy variants of k_recover (tables in a global scratch): quick timing + bitmap check
mkdir -p gpurun_out
: > gpurun_out/quick_r02_occ.jsonl
for v in default gtab3 gtab4 gtab5; do
  if [ $v = default ]; then unset IBFT_LIB; else export IBFT_LIB=$PWD/go-ibft_b200/variants/lib_$v.so; fi
  timeout 300 python tools/quick_bench.py 20 >> gpurun_out/quick_r02_occ.jsonl 2>> gpurun_out/quick_r02_occ.err
done
cat gpurun_out/quick_r02_occ.jsonl
