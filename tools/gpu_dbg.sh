#!/bin/bash
# last check of the round: default bench, both arms, with the tuned CPU arm in the cpu_baseline / reference legs
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_v7_reference_n1.json 2> gpurun_out/bench_r02_v7_reference_n1.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench_r02_v7_n1.json 2> gpurun_out/bench_r02_v7_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads(open('gpurun_out/bench_r02_v7_reference_n1.json').read().strip().splitlines()[-1]); print(r['value'], r['cpu_baseline'])
d=json.loads(open('gpurun_out/bench_r02_v7_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], json.dumps(d['cpu_baseline'])[:1200]); print(json.dumps(d['quorum_latency_us']['cpu'])[:600])
PY
tail -2 gpurun_out/bench_r02_v7_n1.err
