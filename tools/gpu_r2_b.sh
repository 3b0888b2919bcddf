#!/bin/bash
# round 2, call B: the new bench line (uint8 e2e, gpu_reference, configs) at N=1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/profile_bs64_r2b.json > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_r2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','latency_bs1_ms_p50','launch_shapes','clocks'): print(k, d.get(k))
print('e2e', d['e2e'])
print('roofline frac', d['roofline']['frac'], d['roofline']['per_op_back_to_back'])
print('gpu_reference', json.dumps(d.get('gpu_reference'))[:1500])
print('configs', json.dumps(d.get('configs'))[:3000])
print('strong', d.get('strong_bs64'))
PY
cp coco-dataset-based-light-weight-fast-object-detection-model_b200/tune_cache.json gpurun_out/tune_cache_r2b.json 2>/dev/null
