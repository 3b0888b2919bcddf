#!/bin/bash
# 2-GPU functional + throughput check of the NMS-fused peer gather (run with gpurun --gpus 2).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post.py -x -q > gpurun_out/multi_post_tests.log 2>&1
echo "post tests rc=$?"; tail -3 gpurun_out/multi_post_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  tools/dist_check.py > gpurun_out/multi_dist_check.log 2>&1
echo "dist_check rc=$?"; grep dist_check gpurun_out/multi_dist_check.log; tail -5 gpurun_out/multi_dist_check.log
for pg in 1 0; do
  YX_PEER_GATHER=$pg timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/multi_bench_pg$pg.json 2> gpurun_out/multi_bench_pg$pg.err
  echo "bench pg=$pg rc=$?"; tail -c 600 gpurun_out/multi_bench_pg$pg.json; tail -3 gpurun_out/multi_bench_pg$pg.err
done
