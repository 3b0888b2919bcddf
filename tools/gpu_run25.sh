#!/bin/bash
# round-1 evidence run: full default bench (with CPU baseline), reference arm, ncu launch list with DRAM traffic,
# ncu --set full of the top conv launches + post-processing, bs1 profile
mkdir -p gpurun_out
LOG=gpurun_out/run25.log
: > $LOG
timeout 900 python bench.py --profile-out gpurun_out/profile_bs64.json > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo "bench exit=$?" >> $LOG
cat gpurun_out/bench25.json >> $LOG
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench25_ref.json 2>> gpurun_out/bench25.err; echo "ref exit=$?" >> $LOG
cat gpurun_out/bench25_ref.json >> $LOG
python tools/profile_ops.py 1 1280 gpurun_out/profile_bs1.json 5 >> $LOG 2>&1
YX_STEPS=2 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -c 400 --csv --log-file gpurun_out/launches_step.csv python tools/ncu_target.py >> $LOG 2>&1
echo "ncu list exit=$?" >> $LOG
# --set full of individual launches of the profiled step (kernel instantiation = <ACT, RES, MODE, PAIR>)
cap() {  # name, kernel regex, skip, count
  YX_STEPS=2 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu25_$1 python tools/ncu_target.py > gpurun_out/ncu25_$1.log 2>&1
  echo "ncu $1 exit=$?" >> $LOG
}
cap pairhalo "conv_gemm_kernel<2, 0, 1, true>" 4 3
cap pairgeneric "conv_gemm_kernel<2, 0, 0, true>" 0 2
cap generic "conv_gemm_kernel<2, 0, 0, false>" 0 3
cap inplace "conv_gemm_kernel<2, 2" 0 2
cap post "select_infer|sort_keys|nms_kernel|s2d|spp" 0 5
grep -E "exit=|sum ops" $LOG
cut -c1-700 gpurun_out/bench25.json
