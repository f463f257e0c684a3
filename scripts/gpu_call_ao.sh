#!/bin/bash
# rows per warp = line length for any line length (short last slice, realigned row_ptr windows): full GPU suite, then 250^3 / 252^3 / 256^3
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_ao.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_ao.log
tail -3 gpurun_out/gpu_tests_ao.log
timeout 600 python tools/sweep.py --workload laplace3d_250_n128_f64 --algo vector --steps 10 --points ";rows_per_warp=256;" --out gpurun_out/sweepao_l3d250_n128.jsonl > gpurun_out/sweepao_l3d250_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_252_n128_f64 --algo vector --steps 10 --points ";rows_per_slice=16;" --out gpurun_out/sweepao_l3d252_n128.jsonl > gpurun_out/sweepao_l3d252_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points ";;" --out gpurun_out/sweepao_l3d_n128.jsonl > gpurun_out/sweepao_l3d_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points ";;" --out gpurun_out/sweepao_l3d_n64.jsonl > gpurun_out/sweepao_l3d_n64.log 2>&1
echo done
