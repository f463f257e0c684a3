#!/bin/bash
# merge-path kernel with grouped lanes (32/G chunks per warp side by side): parity, then same-box A/B on R-MAT f64 / f32;
# plus the remaining slice variants of the grouped vector kernel at 128 columns
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "merge or grouped" > gpurun_out/gpu_tests_x.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_x.log
tail -3 gpurun_out/gpu_tests_x.log
PR64=";lanes_per_row=8,merge_items=96;lanes_per_row=8,merge_items=128;lanes_per_row=8,merge_items=192,warps_per_cta=4;lanes_per_row=16,merge_items=192;lanes_per_row=16,merge_items=384,warps_per_cta=4;lanes_per_row=8,merge_items=64;;lanes_per_row=8,merge_items=96"
timeout 900 python tools/sweep.py --workload rmat20_n64_f64 --algo merge --steps 10 --points "$PR64" --out gpurun_out/sweepx_rmat_f64.jsonl > gpurun_out/sweepx_rmat_f64.log 2>&1
PR32=";lanes_per_row=8,merge_items=128;lanes_per_row=4,merge_items=64;lanes_per_row=4,merge_items=96;lanes_per_row=8,merge_items=192;prefer_wide_rows=-1;;lanes_per_row=4,merge_items=128,warps_per_cta=4"
timeout 900 python tools/sweep.py --workload rmat20_n64_f32 --algo merge --steps 10 --points "$PR32" --out gpurun_out/sweepx_rmat_f32.jsonl > gpurun_out/sweepx_rmat_f32.log 2>&1
P128=";lanes_per_row=16,reg_flavour=7,rows_per_slice=16,stages=2;lanes_per_row=16,reg_flavour=7,rows_per_slice=8,stages=2;reg_flavour=7,rows_per_slice=16,stages=2;;lanes_per_row=16,reg_flavour=7,rows_per_slice=16,stages=2"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points "$P128" --out gpurun_out/sweepx_l3d_n128.jsonl > gpurun_out/sweepx_l3d_n128.log 2>&1
echo done
