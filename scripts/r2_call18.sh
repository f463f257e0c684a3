#!/bin/bash
# Round 2, GPU call 18: two-lane / one-lane grouped shapes (2 or 4 register tiles per lane, 16 / 32 rows side by side, flat streams) on the narrow stencil products
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export BSM_B200_LIB=$PWD/basic_sparse_matrix_b200/lib_exp/libbsm_b200.so
run() { timeout 300 python tools/sweep.py --workload $1 --algo vector --steps 10 --points "$2" --out gpurun_out/r2_sweep_twolanes_$1.jsonl > gpurun_out/r2c18_$1.log 2>&1; echo "$1 rc=$?"; python tools/show_sweep.py gpurun_out/r2_sweep_twolanes_$1.jsonl | cut -c1-175; }
run laplace3d_256_n4_f64 ";lanes_per_row=1;lanes_per_row=1,rows_per_slice=64;lanes_per_row=1,rows_per_slice=128;lanes_per_row=1,reg_flavour=7;lanes_per_row=1,reg_flavour=7,rows_per_slice=64;lanes_per_row=1,rows_per_slice=64,stages=2"
run laplace3d_256_n8_f64 ";lanes_per_row=2;lanes_per_row=2,rows_per_slice=64;lanes_per_row=2,reg_flavour=7;lanes_per_row=2,reg_flavour=7,rows_per_slice=64;lanes_per_row=1;lanes_per_row=1,rows_per_slice=64;lanes_per_row=1,reg_flavour=7,rows_per_slice=64"
run laplace3d_256_n16_f64 ";lanes_per_row=2;lanes_per_row=2,rows_per_slice=64;lanes_per_row=2,reg_flavour=7;lanes_per_row=2,reg_flavour=7,rows_per_slice=64;rows_per_slice=64,reg_flavour=7"
