#!/bin/bash
# Round 2, GPU call 11: narrow shapes (x4 / x8 f64 on the stencil) as flat entry streams per lane group (a second build,
# -DBSM_EXP_FLAT_NARROW=1, in lib_exp/) against the row-by-row walk of the shipped build, same box.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
PTS=";rows_per_slice=64;rows_per_slice=128;rows_per_slice=64,stages=2;rows_per_slice=128,stages=2;rows_per_slice=256,stages=2;warps_per_cta=8,rows_per_slice=128"
for wl in laplace3d_256_n8_f64 laplace3d_256_n4_f64 band_1m_hb32_n32_f32; do
  for lib in lib lib_exp; do
    BSM_B200_LIB=$PWD/basic_sparse_matrix_b200/$lib/libbsm_b200.so timeout 300 python tools/sweep.py --workload $wl --algo vector --steps 10 --points "$PTS" --out gpurun_out/r2_sweep_flatnarrow_${wl}_$lib.jsonl > gpurun_out/r2c11_${wl}_$lib.log 2>&1; echo "$wl $lib rc=$?"
    python tools/show_sweep.py gpurun_out/r2_sweep_flatnarrow_${wl}_$lib.jsonl 2>/dev/null | cut -c1-175
  done
done
