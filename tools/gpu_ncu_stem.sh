#!/bin/bash
# ncu --set full (+ source) of the image-fed stem (first conv kernel of a step)
mkdir -p gpurun_out
YX_B=16 YX_STEPS=2 YX_TUNE=0 timeout 300 python tools/ncu_target.py > gpurun_out/ncu_stem_plain.log 2>&1 &&
YX_B=16 YX_STEPS=2 YX_TUNE=0 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm -c 1 -f -o gpurun_out/ncu_stem python tools/ncu_target.py > gpurun_out/ncu_stem.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/ncu_stem.ncu-rep
