#!/bin/bash
# A/B of the N-tile rotation (YX_NROT) on the tuned bench model + conv / model parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_i.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_i.log
for f in 1 0; do
  YX_NROT=$f YX_TUNE_CACHE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --profile-out gpurun_out/profile_nrot$f.json > gpurun_out/bench_nrot$f.json 2> gpurun_out/bench_nrot$f.err; echo "bench nrot=$f rc=$?"
done
python - <<'PY'
import json
for k in (1,0):
    d=json.loads(open(f'gpurun_out/bench_nrot{k}.json').read().strip().splitlines()[-1]); p=json.load(open(f'gpurun_out/profile_nrot{k}.json'))['ops']
    print('nrot',k,'value',d['value'],'ms',d['ms_per_step'],'net',d['roofline']['network_ms_in_step'],'clk',d['clocks']['sm_mhz'],'per-op sum',sum(o['ms'] for o in p), 'bs1',d['latency_bs1_ms_p50'])
    for o in p:
        if '288' in o['shape'] or '1152' in o['shape']: print('   ',o['name'][:40], round(o['ms'],3), o['shape'][-110:])
PY
