#!/bin/bash
# stride detection relative to the diagonal + slices that divide the line: full GPU suite, 2-D Laplacian x64, headline and 1/8 block re-check
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_am.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_am.log
tail -2 gpurun_out/gpu_tests_am.log
timeout 600 python tools/sweep.py --workload laplace2d_4096_n64_f64 --algo vector --steps 10 --points ";rows_per_slice=24;rows_per_slice=32;;lanes_per_row=32" --out gpurun_out/sweepam_l2d_n64.jsonl > gpurun_out/sweepam_l2d_n64.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points ";" --out gpurun_out/sweepam_l3d_n128.jsonl > gpurun_out/sweepam_l3d_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --slice 3/8 --points ";" --out gpurun_out/sweepam_l3d_n128_s8.jsonl > gpurun_out/sweepam_l3d_n128_s8.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --slice 3/8 --points ";" --out gpurun_out/sweepam_l3d_n64_s8.jsonl > gpurun_out/sweepam_l3d_n64_s8.log 2>&1
timeout 900 python tools/probe_near_diag.py > gpurun_out/probe_near_diag.jsonl 2> gpurun_out/probe_near_diag.err
echo done
