#!/bin/bash
# 24-warp single-CTA flavour experiment (same box A/B against defaults)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tuning or kats or random_shape" > gpurun_out/gpu_tests_q.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_q.log
tail -3 gpurun_out/gpu_tests_q.log
PV=";reg_flavour=6;reg_flavour=7;reg_flavour=6,rows_per_slice=16;reg_flavour=7,rows_per_slice=32;reg_flavour=6,warps_per_cta=20;reg_flavour=7,warps_per_cta=20;reg_flavour=6,stages=2;reg_flavour=7,stages=2;"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points "$PV" --out gpurun_out/sweepq_l3d_n128.jsonl > gpurun_out/sweepq_l3d_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$PV" --out gpurun_out/sweepq_l3d_n64.jsonl > gpurun_out/sweepq_l3d_n64.log 2>&1
echo done
