#!/bin/bash
# role ablation of the first layers (YX_CONV_DIAG: 1 no epilogue, 2 no A loads, 4 no MMAs), heuristic shapes
mkdir -p gpurun_out
for d in 0 1 4 5; do
  YX_CONV_DIAG=$d YX_TUNE=0 YX_TUNE_CACHE=0 timeout 300 python tools/profile_ops.py 64 1280 gpurun_out/prof_diag$d.json 5 > /dev/null 2>&1
  python - <<PY
import json
o=json.load(open('gpurun_out/prof_diag$d.json'))['ops']
print('diag $d:', ' | '.join(f"{x['name'].split('.')[-2] if x['name'].count('.')>1 else x['name']}.{x['name'].split('.')[-1]} {x['ms']:.3f}" for x in o[:8]))
print('      ', o[0]['shape'][-100:])
PY
done
