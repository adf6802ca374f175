#!/bin/bash
# exactly what the driver runs at N > 1: the DEFAULT bench under torchrun, both arms
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_r02_v6_reference_n$N.json 2> gpurun_out/bench_r02_v6_reference_n$N.err
echo "ref rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r02_v6_default_n$N.json 2> gpurun_out/bench_r02_v6_default_n$N.err
echo "ours rc=$?"; wc -l gpurun_out/bench_r02_v6_default_n$N.json gpurun_out/bench_r02_v6_reference_n$N.json; head -c 400 gpurun_out/bench_r02_v6_default_n$N.json; echo; head -c 300 gpurun_out/bench_r02_v6_reference_n$N.json; echo; tail -3 gpurun_out/bench_r02_v6_default_n$N.err
