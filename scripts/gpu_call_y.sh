#!/bin/bash
# full GPU suite on the new defaults; refinement sweeps (merge items, 16-warp deep-window grouped variant); full default bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_y.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_y.log
tail -3 gpurun_out/gpu_tests_y.log
PR64=";merge_items=384,lanes_per_row=32;lanes_per_row=16,merge_items=256;lanes_per_row=16,merge_items=160;lanes_per_row=16,merge_items=224;;merge_items=384,lanes_per_row=32"
timeout 900 python tools/sweep.py --workload rmat20_n64_f64 --algo merge --steps 10 --points "$PR64" --out gpurun_out/sweepy_rmat_f64.jsonl > gpurun_out/sweepy_rmat_f64.log 2>&1
PR32=";merge_items=384,lanes_per_row=32;lanes_per_row=8,merge_items=96;lanes_per_row=8,merge_items=160;lanes_per_row=16,merge_items=192;;merge_items=384,lanes_per_row=32"
timeout 900 python tools/sweep.py --workload rmat20_n64_f32 --algo merge --steps 10 --points "$PR32" --out gpurun_out/sweepy_rmat_f32.jsonl > gpurun_out/sweepy_rmat_f32.log 2>&1
P64=";lanes_per_row=8,reg_flavour=1,rows_per_slice=16,stages=2;lanes_per_row=8,reg_flavour=1,rows_per_slice=32,stages=2;;lanes_per_row=8,reg_flavour=1,rows_per_slice=16,stages=3"
timeout 900 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$P64" --out gpurun_out/sweepy_l3d_n64.jsonl > gpurun_out/sweepy_l3d_n64.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err
tail -c 300 gpurun_out/bench_y.json
echo done
