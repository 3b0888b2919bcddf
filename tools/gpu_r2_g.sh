#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for f in 0 1; do
  YX_FUSE_S2D=$f YX_TUNE_CACHE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --profile-out gpurun_out/profile_fuse$f.json > gpurun_out/bench_fuse$f.json 2> gpurun_out/bench_fuse$f.err; echo "bench fuse=$f rc=$?"; tail -2 gpurun_out/bench_fuse$f.err
done
python - <<'PY'
import json
for k in (0,1):
    d=json.loads(open(f'gpurun_out/bench_fuse{k}.json').read().strip().splitlines()[-1]); p=json.load(open(f'gpurun_out/profile_fuse{k}.json'))['ops']
    print('fuse',k,'value',d['value'],'ms',d['ms_per_step'],'net',d['roofline']['network_ms_in_step'],'clk',d['clocks']['sm_mhz'],'per-op sum',sum(o['ms'] for o in p), 'bs1',d['latency_bs1_ms_p50'])
    for o in p[:3]: print('   ',o['name'][:40], round(o['ms'],3), o['shape'][-100:])
PY
