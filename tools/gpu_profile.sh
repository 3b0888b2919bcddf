#!/bin/bash
# ncu evidence (B200_PROFILING.md recipe): launch list of one step + full captures of the two top conv launches.
mkdir -p gpurun_out
LOG=gpurun_out/profile.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> $LOG
tail -8 gpurun_out/pytest_gpu.log >> $LOG
python tools/ncu_target.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 128 -c 128 --csv --log-file gpurun_out/launches.csv \
    python tools/ncu_target.py >> $LOG 2>&1
echo "launch list exit=$?" >> $LOG
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 217 -c 1 -o gpurun_out/prof_head python tools/ncu_target.py >> $LOG 2>&1
echo "ncu head exit=$?" >> $LOG
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 120 -c 1 -o gpurun_out/prof_stem python tools/ncu_target.py >> $LOG 2>&1
echo "ncu stem exit=$?" >> $LOG
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 123 -c 1 -o gpurun_out/prof_1x1 python tools/ncu_target.py >> $LOG 2>&1
echo "ncu 1x1 exit=$?" >> $LOG
ncu --set full --clock-control none -k regex:"select_infer|sort_keys|nms_kernel" -s 3 -c 3 -o gpurun_out/prof_post python tools/ncu_target.py >> $LOG 2>&1
echo "ncu post exit=$?" >> $LOG
ls -la gpurun_out >> $LOG
grep -E "exit=|passed|failed|FAILED" $LOG
