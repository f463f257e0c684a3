#!/bin/bash
# full GPU suite (new randomized sweep, checksum-of-checksums, C++ mul_dense_s) on one GPU
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_n.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_n.log
tail -5 gpurun_out/gpu_tests_n.log
echo done
