#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run10.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> $LOG
tail -6 gpurun_out/pytest_gpu.log >> $LOG
timeout 300 python tools/profile_ops.py 1 1280 gpurun_out/profile_bs1.json 20 >> $LOG 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline >> $LOG 2>&1 || echo "bench exit=$?" >> $LOG
grep -E "exit=|passed|failed|FAILED|sum ops|value" $LOG | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_last.json')) if False else None
PY
grep -o '"latency_bs1_ms_p50": [0-9.]*' $LOG
