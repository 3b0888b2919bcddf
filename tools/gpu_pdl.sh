#!/bin/bash
# A/B of programmatic dependent launch on the bench step, then the whole GPU test suite with it on.
mkdir -p gpurun_out
for pdl in 1 0 1 0; do
  YX_PDL=$pdl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err
  echo "pdl=$pdl rc=$? $(python -c "import json;j=json.loads([l for l in open('gpurun_out/bench_pdl$pdl.json') if l.startswith('{')][-1]);print(j['value'], j['ms_per_step'], j['e2e']['value'], j['latency_bs1_ms_p50'], j['clocks']['sm_mhz'])")"
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit=$?"
tail -4 gpurun_out/pytest_gpu.log
