#!/bin/bash
# f32 x128 (512-byte rows, the f32 twin of the north_star shape): grouped default vs a warp per row
set -u
mkdir -p gpurun_out
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f32 --algo vector --steps 10 --points ";lanes_per_row=32;;lanes_per_row=32;lanes_per_row=16;lanes_per_row=8,rows_per_slice=32" --out gpurun_out/sweepae_l3d_n128_f32.jsonl > gpurun_out/sweepae_l3d_n128_f32.log 2>&1
echo done
