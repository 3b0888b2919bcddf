#!/bin/bash
# N-GPU functional + throughput check of the NMS-fused peer gather (run with gpurun --gpus N; N from $1, default 2).
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/dist_check.py > gpurun_out/multi_dist_check_n$N.log 2>&1
echo "dist_check rc=$?"; grep dist_check gpurun_out/multi_dist_check_n$N.log | sort | uniq -c | head -40
for pg in 1 0; do
  YX_PEER_GATHER=$pg timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/multi_bench_n${N}_pg$pg.json 2> gpurun_out/multi_bench_n${N}_pg$pg.err
  echo "bench pg=$pg rc=$?"; cut -c1-330 gpurun_out/multi_bench_n${N}_pg$pg.json; tail -3 gpurun_out/multi_bench_n${N}_pg$pg.err | cut -c1-300
done
