#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/perf_reference_gpu.py 64 1280 5 > gpurun_out/ref_gpu_bs64.json 2> gpurun_out/ref_gpu.err || echo "ref exit=$?"
timeout 600 python tests/perf_reference_gpu.py 1 1280 30 > gpurun_out/ref_gpu_bs1.json 2>> gpurun_out/ref_gpu.err || echo "ref1 exit=$?"
cat gpurun_out/ref_gpu_bs64.json gpurun_out/ref_gpu_bs1.json; tail -3 gpurun_out/ref_gpu.err
