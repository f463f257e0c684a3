#!/bin/bash
# Round 2, GPU call 24: B rows far from the slice's rows (the +-plane neighbours, read once per SM) loaded with L1::no_allocate (second build), same box A/B
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for wl in laplace3d_256_n128_f64 laplace3d_256_n64_f64 laplace2d_4096_n64_f64; do
  for lib in lib lib_exp lib lib_exp; do
    BSM_B200_LIB=$PWD/basic_sparse_matrix_b200/$lib/libbsm_b200.so timeout 300 python tools/sweep.py --workload $wl --steps 20 --points ";" --out gpurun_out/r2_sweep_far_noalloc_${wl}_$lib.jsonl > gpurun_out/r2c24.log 2>&1
    python tools/show_sweep.py gpurun_out/r2_sweep_far_noalloc_${wl}_$lib.jsonl | tail -1 | cut -c1-100 | sed "s/^/$wl $lib /"
  done
done
