#!/bin/bash
# Round 2, GPU call 25: last sweep of the narrow shapes' launch parameters (x16 / x4 / x8 f64 on the stencil)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { timeout 300 python tools/sweep.py --workload $1 --algo vector --steps 20 --points "$2" --out gpurun_out/r2_sweep_narrow_final_$1.jsonl > gpurun_out/r2c25_$1.log 2>&1; echo "$1 rc=$?"; python tools/show_sweep.py gpurun_out/r2_sweep_narrow_final_$1.jsonl | cut -c1-170; }
run laplace3d_256_n16_f64 ";rows_per_slice=16;rows_per_slice=24;rows_per_slice=48;stages=3;rows_per_slice=16,stages=3;reg_flavour=7;reg_flavour=7,rows_per_slice=16;reg_flavour=7,rows_per_slice=16,stages=3;rows_per_warp=128;rows_per_warp=64;rows_per_warp=512;ctas_per_sm=2;warps_per_cta=4"
run laplace3d_256_n4_f64 ";rows_per_slice=32;rows_per_slice=32,stages=3;rows_per_slice=48,stages=2;rows_per_slice=96,stages=1;rows_per_warp=128;rows_per_warp=64;warps_per_cta=12;warps_per_cta=8;rows_per_slice=32,stages=4"
run laplace3d_256_n8_f64 ";rows_per_warp=128;rows_per_warp=64;rows_per_slice=64,stages=2,warps_per_cta=12;rows_per_slice=48;rows_per_slice=80"
