#!/bin/bash
# sparse-MMA probe, one experiment per process (tools/sp_probe.cu)
mkdir -p gpurun_out
OUT=gpurun_out/sp_probe.txt
: > $OUT
for t in "0 0 0" "0 0 1" "0 0 2" "0 0 3" "0 2 0" "0 2 1" "0 1 0" "1 0 0" "2 0 0" "3 0 0" "2 0 1" "3 0 1"; do
  timeout 60 ./tools/sp_probe decode $t >> $OUT 2>&1 || echo "decode $t FAILED rc=$?" >> $OUT
done
for pair in 0 1; do for sp in 0 1; do for n in 64 128 192 256; do
  timeout 60 ./tools/sp_probe time $pair $sp $n >> $OUT 2>&1 || echo "time $pair $sp $n FAILED rc=$?" >> $OUT
done; done; done
grep -c "FAILED" $OUT; grep "FAILED\|^time\|CUDA error" $OUT | head -40
