#!/bin/bash
# conv + model parity tests, then a short bench with the per-op profile
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_bs64.json > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit=$?"
python -c "
import json
j=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['latency_bs1_ms_p50'], j['clocks']['sm_mhz'], j['roofline']['frac'], j['roofline']['per_op_back_to_back']['frac'])
o=json.load(open('gpurun_out/profile_bs64.json'))['ops']
print('per-op sum', sum(x['ms'] for x in o), 'stem', o[1]['ms'], 'dark2.1.conv1+2', o[3]['ms'])"
