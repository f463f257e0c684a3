#!/bin/bash
# line lengths that are not powers of two: 252 (divisible by 4: slices of 12 rows) and 250 (not divisible: rows per warp rounded up)
set -u
mkdir -p gpurun_out
timeout 600 python tools/sweep.py --workload laplace3d_252_n128_f64 --algo vector --steps 10 --points ";rows_per_slice=16;rows_per_slice=28;rows_per_slice=36" --out gpurun_out/sweepan_l3d252_n128.jsonl > gpurun_out/sweepan_l3d252_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_252_n64_f64 --algo vector --steps 10 --points ";rows_per_slice=16;rows_per_slice=28;rows_per_slice=36" --out gpurun_out/sweepan_l3d252_n64.jsonl > gpurun_out/sweepan_l3d252_n64.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_250_n128_f64 --algo vector --steps 10 --points ";rows_per_warp=16;rows_per_warp=128;rows_per_warp=500" --out gpurun_out/sweepan_l3d250_n128.jsonl > gpurun_out/sweepan_l3d250_n128.log 2>&1
echo done
