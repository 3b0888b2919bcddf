#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/second.log
: > $LOG
timeout 600 python tools/gpu_debug.py drill yolox_m_p6 640 640 2 >> $LOG 2>&1 || echo "drill exit=$?" >> $LOG
timeout 900 python -m pytest tests -m gpu -x -q >> gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> $LOG
tail -15 gpurun_out/pytest_gpu.log >> $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1 || echo "smoke exit=$?" >> $LOG
timeout 900 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/profile_bs64.json > gpurun_out/bench.json 2>> $LOG || echo "bench exit=$?" >> $LOG
cat gpurun_out/bench.json >> $LOG
tail -40 $LOG
