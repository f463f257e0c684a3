#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gpu_tests_r.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_r.log
tail -15 gpurun_out/gpu_tests_r.log
echo done
