#!/bin/bash
# strong + weak scaling at N ranks of one node (development run: --skip-extras)
N=$1
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --skip-extras --latency-reps 200 > gpurun_out/bench_r02_v3_n$N.json 2> gpurun_out/bench_r02_v3_n$N.err
echo "rc=$?"; tail -c 2500 gpurun_out/bench_r02_v3_n$N.json; tail -3 gpurun_out/bench_r02_v3_n$N.err
