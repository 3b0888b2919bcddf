#!/bin/bash
# the two captures the round-2 evidence run missed (sparse, depthwise), the sparse-vs-dense per-op table, the dw table
mkdir -p gpurun_out
LOG=gpurun_out/evidence_r02b.log
: > $LOG
cap() {
  env $5 YX_STEPS=2 timeout 600 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu_r02_$1 python tools/ncu_target.py > gpurun_out/ncu_r02_$1.log 2>&1
  echo "ncu $1 exit=$? $(ls -la gpurun_out/ncu_r02_$1.ncu-rep 2>/dev/null | awk '{print $5}')" >> $LOG
  ncu -i gpurun_out/ncu_r02_$1.ncu-rep --page raw --csv > gpurun_out/ncu_r02_$1_raw.csv 2>/dev/null
  rm -f gpurun_out/ncu_r02_$1.ncu-rep
}
timeout 900 python tools/sparse_profile.py 64 1280 gpurun_out/sparse_profile_r02.md >> $LOG 2>&1; echo "sparse profile exit=$?" >> $LOG
YX_MASKS=two_four YX_SPARSE=force YX_STEPS=2 timeout 600 python tools/ncu_target.py > gpurun_out/ncu_plain_sparse_r02.log 2>&1 &&
cap sparse "conv_gemm_kernel<.int.2, .int.[02], .int.1, .bool.0, .bool.1, .bool.0>" 8 1 "YX_B=64 YX_MASKS=two_four YX_SPARSE=force"
YX_MODEL=dw YX_B=32 YX_S=640 YX_STEPS=2 timeout 600 python tools/ncu_target.py > gpurun_out/ncu_plain_dw_r02.log 2>&1 &&
cap dwconv "dwconv_tile_kernel" 3 2 "YX_B=32 YX_S=640 YX_MODEL=dw"
python tools/dw_profile.py 32 640 > gpurun_out/dw_profile_r02.txt 2>&1
YX_DW_TILE=0 python tools/dw_profile.py 32 640 >> gpurun_out/dw_profile_r02.txt 2>&1
cat $LOG | tail -8; head -3 gpurun_out/dw_profile_r02.txt | cut -c1-200; head -4 gpurun_out/sparse_profile_r02.md | cut -c1-300
