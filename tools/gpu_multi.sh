#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/multi.log
: > $LOG
nvidia-smi -L >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline >> $LOG 2>&1 || echo "bench2 exit=$?" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 >> $LOG 2>&1 || echo "ref2 exit=$?" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_check.py >> $LOG 2>&1 || echo "distcheck exit=$?" >> $LOG
grep -E "exit=|value|dist_check|GPU " $LOG | cut -c1-700
