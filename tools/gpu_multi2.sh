#!/bin/bash
# N-GPU functional + throughput check, round 2 (pipelined 3-window gather, strong_bs64)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/dist_check.py > gpurun_out/multi2_dist_check_n$N.log 2>&1
echo "dist_check rc=$?"; grep dist_check gpurun_out/multi2_dist_check_n$N.log | sort | uniq -c | head -40; tail -5 gpurun_out/multi2_dist_check_n$N.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/multi2_bench_n${N}.json 2> gpurun_out/multi2_bench_n${N}.err
echo "bench rc=$?"; tail -3 gpurun_out/multi2_bench_n${N}.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/multi2_bench_n${N}.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'net ms',d['roofline']['network_ms_in_step'],'e2e',d['e2e']['value'])
print('strong',d.get('strong_bs64')); print('gather',d.get('gather'))
PY
