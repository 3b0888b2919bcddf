#!/bin/bash
# one more --set full capture: a CTA-pair generic (streamed 1x1) layer of the bench step
mkdir -p gpurun_out
YX_B=64 YX_STEPS=2 timeout 600 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled \
  -k regex:"conv_gemm_kernel<.int.2, .int.0, .int.0, .bool.1, .bool.0, .bool.0>" -s 20 -c 3 -f -o gpurun_out/ncu_r02_pairgeneric python tools/ncu_target.py > gpurun_out/ncu_r02_pairgeneric.log 2>&1
echo "exit=$?"
ncu -i gpurun_out/ncu_r02_pairgeneric.ncu-rep --page raw --csv > gpurun_out/ncu_r02_pairgeneric_raw.csv 2>/dev/null
rm -f gpurun_out/ncu_r02_pairgeneric.ncu-rep
wc -c gpurun_out/ncu_r02_pairgeneric_raw.csv
