#!/bin/bash
# Round-2 evidence run (one GPU).  Everything lands in gpurun_out/ with "r02" in the name;
# `python tools/make_profiles.py r02 r02` turns it into the committed summaries under profiles/.
#   1. default bench line (full: gpu_reference, configs, cpu baseline) + reference arm + per-op profile bs64, per-op bs1
#   2. ncu launch list of ONE steady-state step (times + DRAM bytes per launch)
#   3. ncu --set full captures of the kernels the roofline numbers are about
mkdir -p gpurun_out
LOG=gpurun_out/evidence_r02.log
: > $LOG
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/profile_bs64_r02.json > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench exit=$?" >> $LOG
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02_ref.json 2>> gpurun_out/bench_r02_n1.err; echo "ref exit=$?" >> $LOG
cp coco-dataset-based-light-weight-fast-object-detection-model_b200/tune_cache.json gpurun_out/tune_cache_r02.json 2>/dev/null
python tools/profile_ops.py 1 1280 gpurun_out/profile_bs1_r02.json 5 >> $LOG 2>&1
YX_STEPS=2 timeout 600 python tools/ncu_target.py > gpurun_out/ncu_plain_r02.log 2>&1 &&
YX_STEPS=2 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -c 400 --csv --log-file gpurun_out/launches_step_r02.csv python tools/ncu_target.py >> $LOG 2>&1
echo "ncu list exit=$?" >> $LOG
cap() {  # name, demangled kernel regex, skip, count, extra env
  env $5 YX_STEPS=2 timeout 600 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu_r02_$1 python tools/ncu_target.py > gpurun_out/ncu_r02_$1.log 2>&1
  echo "ncu $1 exit=$? $(ls -la gpurun_out/ncu_r02_$1.ncu-rep 2>/dev/null | awk '{print $5}')" >> $LOG
  ncu -i gpurun_out/ncu_r02_$1.ncu-rep --page raw --csv > gpurun_out/ncu_r02_$1_raw.csv 2>/dev/null
  rm -f gpurun_out/ncu_r02_$1.ncu-rep   # the raw page is what the summaries are built from; reports exceed the 64 MiB return limit
}
if [ -z "$YX_SKIP_CAPTURES" ]; then
cap pairhalo "conv_gemm_kernel<.int.2, .int.0, .int.1, .bool.1, .bool.0, .bool.0>" 5 1 YX_B=64
cap imagestem "conv_gemm_kernel<.int.2, .int.0, .int.[12], .bool.0, .bool.0, .bool.1>" 0 1 YX_B=64
cap generic "conv_gemm_kernel<.int.2, .int.0, .int.0, .bool.0, .bool.0, .bool.0>" 0 1 YX_B=64
cap post "select_infer|sort_keys|nms_kernel|spp_" 0 4 YX_B=64
YX_MASKS=two_four YX_SPARSE=force YX_STEPS=2 timeout 600 python tools/ncu_target.py > gpurun_out/ncu_plain_sparse_r02.log 2>&1 &&
cap sparse "conv_gemm_kernel<.int.2, .int.[02], .int.1, .bool.0, .bool.1, .bool.0>" 8 1 "YX_B=64 YX_MASKS=two_four YX_SPARSE=force"
YX_MODEL=dw YX_B=32 YX_S=640 YX_STEPS=2 timeout 600 python tools/ncu_target.py > gpurun_out/ncu_plain_dw_r02.log 2>&1 &&
cap dwconv "dwconv_tile_kernel" 3 2 "YX_B=32 YX_S=640 YX_MODEL=dw"
fi
grep -E "exit=|sum ops" $LOG
cut -c1-300 gpurun_out/bench_r02_n1.json
