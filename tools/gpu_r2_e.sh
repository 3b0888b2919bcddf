#!/bin/bash
# A/B of "warp 2 joins the A producers" (YX_W2A) on the tuned bench engine + conv parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -3
for w2 in 0 1; do
  YX_W2A=$w2 YX_TUNE_CACHE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --profile-out gpurun_out/profile_w2a$w2.json > gpurun_out/bench_w2a$w2.json 2> gpurun_out/bench_w2a$w2.err; echo "bench w2a=$w2 rc=$?"
done
python - <<'PY'
import json
a=json.load(open('gpurun_out/profile_w2a0.json'))['ops']; b=json.load(open('gpurun_out/profile_w2a1.json'))['ops']
print('per-op sum w2a=0', sum(o['ms'] for o in a), 'w2a=1', sum(o['ms'] for o in b))
for k in (0,1):
    d=json.loads(open(f'gpurun_out/bench_w2a{k}.json').read().strip().splitlines()[-1]); print('w2a',k,'value',d['value'],'ms',d['ms_per_step'],d['clocks']['sm_mhz'])
for i,(x,y) in enumerate(zip(a,b)):
    if abs(x['ms']-y['ms'])>0.015: print(i, x['name'][:36], f"{x['ms']:.3f} -> {y['ms']:.3f}", y['shape'].split(': ')[-1][:90])
PY
