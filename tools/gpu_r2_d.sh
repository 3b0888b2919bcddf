#!/bin/bash
# conv parity (all forced shapes) + bench with per-op profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -8 > gpurun_out/pytest_conv_r2d.log; echo "conv rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/pytest_conv_r2d.log
YX_TUNE_CACHE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --profile-out gpurun_out/profile_bs64_r2d.json > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2d.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','latency_bs1_ms_p50','launch_shapes','clocks'): print(k, d.get(k))
print('roofline frac', d['roofline']['frac'], d['roofline']['per_op_back_to_back'])
p=json.load(open('gpurun_out/profile_bs64_r2d.json'))
print('per-op sum', sum(o['ms'] for o in p['ops']), 'alt picked:', sum('alt' in o['shape'] for o in p['ops']))
PY
