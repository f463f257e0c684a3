#!/bin/bash
# Round 2, GPU call 12: flat entry streams for narrow shapes, second sweep: which lane shapes and slice lengths profit (x8 / x16 / x32 f64,
# shipped build = row by row or grouped; lib_exp = flat streams for every G > 1).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { # workload points
  for lib in lib lib_exp; do
    BSM_B200_LIB=$PWD/basic_sparse_matrix_b200/$lib/libbsm_b200.so timeout 300 python tools/sweep.py --workload $1 --algo vector --steps 10 --points "$2" --out gpurun_out/r2_sweep_flatnarrow2_$1_$lib.jsonl > gpurun_out/r2c12_$1_$lib.log 2>&1; echo "$1 $lib rc=$?"
    python tools/show_sweep.py gpurun_out/r2_sweep_flatnarrow2_$1_$lib.jsonl 2>/dev/null | cut -c1-175
  done
}
run laplace3d_256_n8_f64 ";rows_per_slice=48;rows_per_slice=64;rows_per_slice=96;rows_per_slice=64,warps_per_cta=12;rows_per_slice=96,warps_per_cta=12;rows_per_slice=64,rows_per_warp=64;rows_per_slice=64,rows_per_warp=128;rows_per_slice=32,stages=4"
run laplace3d_256_n16_f64 ";rows_per_slice=32;rows_per_slice=64;lanes_per_row=8;lanes_per_row=8,rows_per_slice=32;lanes_per_row=8,rows_per_slice=64"
run laplace3d_256_n32_f64 ";rows_per_slice=32;lanes_per_row=16;lanes_per_row=16,rows_per_slice=32;lanes_per_row=16,rows_per_slice=64"
run laplace2d_2048_n1_f64 ";"
