#!/bin/bash
# SpMV / very narrow products on the headline matrix: defaults vs rows-per-warp and slice variants
set -u
mkdir -p gpurun_out
timeout 600 python tools/sweep.py --workload laplace3d_256_n1_f64 --algo vector --steps 10 --points ";rows_per_warp=32;rows_per_warp=128;rows_per_warp=1024;rows_per_slice=64;rows_per_slice=256;warps_per_cta=8;" --out gpurun_out/sweepaq_l3d_n1.jsonl > gpurun_out/sweepaq_l3d_n1.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n4_f64 --algo vector --steps 10 --points ";rows_per_warp=32;rows_per_warp=128;rows_per_warp=1024;rows_per_slice=64;" --out gpurun_out/sweepaq_l3d_n4.jsonl > gpurun_out/sweepaq_l3d_n4.log 2>&1
echo done
