#!/bin/bash
# neighbour-row reuse (reg_flavour 9 / 10): parity, then same-box interleaved A/B against the defaults
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reuse or tuning or kats" > gpurun_out/gpu_tests_t.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_t.log
tail -3 gpurun_out/gpu_tests_t.log
P64=";reg_flavour=9;reg_flavour=10;;reg_flavour=9;reg_flavour=10;reg_flavour=9,rows_per_slice=16;reg_flavour=10,rows_per_slice=16;reg_flavour=10,rows_per_slice=16,stages=3;;reg_flavour=9;reg_flavour=10"
timeout 900 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$P64" --out gpurun_out/sweept_l3d_n64.jsonl > gpurun_out/sweept_l3d_n64.log 2>&1
P128=";reg_flavour=9;reg_flavour=10;;reg_flavour=9;reg_flavour=10;reg_flavour=10,rows_per_slice=32,stages=2;reg_flavour=9,rows_per_slice=32,stages=2;;reg_flavour=9;reg_flavour=10"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points "$P128" --out gpurun_out/sweept_l3d_n128.jsonl > gpurun_out/sweept_l3d_n128.log 2>&1
echo done
