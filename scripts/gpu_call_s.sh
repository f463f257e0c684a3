#!/bin/bash
# deeper gather windows within the 85-register budget + slice/stage variants of the 24-warp CTA (same box, interleaved)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tuning or kats or random_shape" > gpurun_out/gpu_tests_s.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_s.log
tail -3 gpurun_out/gpu_tests_s.log
P64=";reg_flavour=8;;reg_flavour=8;reg_flavour=7,rows_per_slice=16,stages=2;reg_flavour=7,rows_per_slice=16,stages=3;reg_flavour=7,rows_per_slice=24,stages=2;reg_flavour=8,rows_per_slice=16,stages=3;reg_flavour=8,rows_per_slice=16,stages=2;;reg_flavour=8"
timeout 900 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$P64" --out gpurun_out/sweeps_l3d_n64.jsonl > gpurun_out/sweeps_l3d_n64.log 2>&1
P128=";reg_flavour=8;;reg_flavour=8;;reg_flavour=8;reg_flavour=8,stages=2;reg_flavour=8,rows_per_slice=32"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points "$P128" --out gpurun_out/sweeps_l3d_n128.jsonl > gpurun_out/sweeps_l3d_n128.log 2>&1
echo done
