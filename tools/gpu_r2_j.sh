#!/bin/bash
# whole -m gpu suite, then the tuned bench (no extras) with the per-op profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_j.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_j.log
YX_TUNE_CACHE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --profile-out gpurun_out/profile_j.json > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1]); p=json.load(open('gpurun_out/profile_j.json'))['ops']
print('value',d['value'],'ms',d['ms_per_step'],'net',d['roofline']['network_ms_in_step'],'clk',d['clocks']['sm_mhz'],'per-op sum',sum(o['ms'] for o in p), 'bs1',d['latency_bs1_ms_p50'], 'frac', d['roofline']['frac'])
for o in p:
    if o['ms']>0.14: print('   ',o['name'][:40], round(o['ms'],3), o['shape'][-118:])
PY
