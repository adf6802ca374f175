#!/bin/bash
# 2-rank probe of the peer-memory exchange: plain, then under memcheck; then the strong-scaling legs of bench at N=2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/p2p_probe.py > gpurun_out/dbg_p2p_n2.txt 2>&1
echo "rc=$?" >> gpurun_out/dbg_p2p_n2.txt
grep -n "bitmap ok\|rror\|rc=" gpurun_out/dbg_p2p_n2.txt | head -20
timeout 400 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 compute-sanitizer --tool memcheck --print-limit 6 python tools/p2p_probe.py > gpurun_out/dbg_p2p_n2_memcheck.txt 2>&1
echo "rc=$?" >> gpurun_out/dbg_p2p_n2_memcheck.txt
grep -n "Invalid\|=========     at\|Address\|ERROR SUMMARY\|bitmap ok\|rc=" gpurun_out/dbg_p2p_n2_memcheck.txt | head -30
if grep -q "rc=0" gpurun_out/dbg_p2p_n2.txt; then
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_v4_n2.json 2> gpurun_out/bench_r02_v4_n2.err
  echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_v4_n2.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('strong_scaling'), indent=1)[:3000])
PY
fi
