#!/bin/bash
# grouped lanes on narrower output rows (256 / 128 bytes) of the stencil matrix: same-box A/B
set -u
mkdir -p gpurun_out
P=";lanes_per_row=8;lanes_per_row=4;lanes_per_row=8,reg_flavour=7;lanes_per_row=4,reg_flavour=7;lanes_per_row=8,reg_flavour=7,rows_per_slice=32,stages=2;lanes_per_row=4,reg_flavour=7,rows_per_slice=32,stages=2;;lanes_per_row=8,reg_flavour=7,rows_per_slice=16,stages=2;lanes_per_row=4,reg_flavour=7,rows_per_slice=16,stages=2"
timeout 600 python tools/sweep.py --workload laplace3d_256_n32_f64 --algo vector --steps 10 --points "$P" --out gpurun_out/sweepz_l3d_n32_f64.jsonl > gpurun_out/sweepz_l3d_n32_f64.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f32 --algo vector --steps 10 --points "$P" --out gpurun_out/sweepz_l3d_n64_f32.jsonl > gpurun_out/sweepz_l3d_n64_f32.log 2>&1
P16=";lanes_per_row=4;lanes_per_row=4,reg_flavour=7;lanes_per_row=4,reg_flavour=7,rows_per_slice=32,stages=2;;lanes_per_row=4,reg_flavour=7,rows_per_slice=16,stages=2"
timeout 600 python tools/sweep.py --workload laplace3d_256_n16_f64 --algo vector --steps 10 --points "$P16" --out gpurun_out/sweepz_l3d_n16_f64.jsonl > gpurun_out/sweepz_l3d_n16_f64.log 2>&1
echo done
