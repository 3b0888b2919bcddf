#!/bin/bash
mkdir -p gpurun_out
cap() {  # name, demangled kernel regex, skip, count
  YX_STEPS=2 timeout 600 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu25_$1 python tools/ncu_target.py > gpurun_out/ncu25_$1.log 2>&1
  echo "ncu $1 exit=$? $(ls -la gpurun_out/ncu25_$1.ncu-rep 2>/dev/null)"
  ncu -i gpurun_out/ncu25_$1.ncu-rep --page raw --csv > gpurun_out/ncu25_$1_raw.csv 2>/dev/null
  rm -f gpurun_out/ncu25_$1.ncu-rep   # the raw page is what the summaries are built from; reports exceed the 64 MiB return limit
}
cap pairhalo "conv_gemm_kernel<.int.2, .int.0, .int.1, .bool.1>" 5 1
cap pairgeneric "conv_gemm_kernel<.int.2, .int.0, .int.0, .bool.1>" 0 1
cap generic "conv_gemm_kernel<.int.2, .int.0, .int.0, .bool.0>" 0 1
cap inplace "conv_gemm_kernel<.int.2, .int.2, .int.2, .bool.0>" 0 1
cap pairinplace "conv_gemm_kernel<.int.2, .int.2, .int.1, .bool.1>" 0 1
cap post "select_infer|sort_keys|nms_kernel|s2d_kernel|spp_" 0 5
