#!/bin/bash
mkdir -p gpurun_out
(./tools/mma_probe chain
python tools/conv_time.py stem16
python tools/conv_trace.py stem16 2 4 2>&1 | grep -E "CASE|trace:|blocked|^ +(8|9|10|11|12) ") > gpurun_out/m.log 2>&1
cut -c1-230 gpurun_out/m.log
