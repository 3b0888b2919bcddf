#!/bin/bash
# Round evidence run (one GPU): default bench (with the CPU baseline) + reference arm, bs1 per-op profile, ncu launch list
# of ONE steady-state step with measured DRAM bytes, and ncu --set full captures of the main conv shapes + the other kernels.
# Everything lands in gpurun_out/; `python tools/make_profiles.py` turns it into the committed summaries under profiles/.
mkdir -p gpurun_out
LOG=gpurun_out/evidence.log
: > $LOG
timeout 900 python bench.py --profile-out gpurun_out/profile_bs64.json > gpurun_out/bench_evidence.json 2> gpurun_out/bench_evidence.err; echo "bench exit=$?" >> $LOG
if [ -z "$YX_SKIP_CAPTURES" ]; then timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_evidence_ref.json 2>> gpurun_out/bench_evidence.err; echo "ref exit=$?" >> $LOG; fi
python tools/profile_ops.py 1 1280 gpurun_out/profile_bs1.json 5 >> $LOG 2>&1
YX_STEPS=2 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -c 400 --csv --log-file gpurun_out/launches_step.csv python tools/ncu_target.py >> $LOG 2>&1
echo "ncu list exit=$?" >> $LOG
if [ -z "$YX_SKIP_CAPTURES" ]; then bash tools/gpu_ncu.sh >> $LOG 2>&1; fi   # --set full captures: ~5 GPU-minutes
grep -E "exit=|sum ops" $LOG
cut -c1-400 gpurun_out/bench_evidence.json
