#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_post.py tests/test_gpu_model.py -x -q 2>&1 | tail -4
for s in 0 1; do YX_DW_STRIP=$s timeout 300 python tools/dw_profile.py 32 640 2>&1 | tail -25 | tee -a gpurun_out/dw_profile_r2.txt; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2f.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'net',d['roofline']['network_ms_in_step'],'bs1 p50',d['latency_bs1_ms_p50'], d['clocks'])
PY
