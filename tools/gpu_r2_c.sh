#!/bin/bash
# round 2, call C: sparse conv parity (forced shapes), then the engine-level sparse test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "sparse" 2>&1 | tail -30 > gpurun_out/pytest_sparse_conv.log; echo "conv sparse rc=${PIPESTATUS[0]}"
tail -25 gpurun_out/pytest_sparse_conv.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -k "two_four or pruned" 2>&1 | tail -30 > gpurun_out/pytest_sparse_model.log; echo "model sparse rc=${PIPESTATUS[0]}"
tail -25 gpurun_out/pytest_sparse_model.log
