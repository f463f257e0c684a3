#!/bin/bash
# grouped 8x4 default for one-tile shapes: full GPU suite, then same-box A/B of the new default against the old one and neighbours
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_v.log
tail -3 gpurun_out/gpu_tests_v.log
P64=";lanes_per_row=32;;lanes_per_row=32;lanes_per_row=8,reg_flavour=7,rows_per_slice=16,stages=3;lanes_per_row=8,reg_flavour=7,rows_per_slice=16,stages=2;lanes_per_row=16,reg_flavour=7,rows_per_slice=32,stages=2;lanes_per_row=8,reg_flavour=5,rows_per_slice=32,stages=2;lanes_per_row=8,reg_flavour=7,rows_per_slice=32,stages=2,rows_per_warp=128;;lanes_per_row=32"
timeout 900 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$P64" --out gpurun_out/sweepv_l3d_n64.jsonl > gpurun_out/sweepv_l3d_n64.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --slice 3/8 --points ";lanes_per_row=32;;lanes_per_row=32;rows_per_warp=128;rows_per_warp=64" --out gpurun_out/sweepv_l3d_n64_s8.jsonl > gpurun_out/sweepv_l3d_n64_s8.log 2>&1
timeout 900 python bench.py --workload laplace3d_256_n64_f64 --steps 30 --warmup 5 --no-extras --no-cpu > gpurun_out/bench_v_n64.json 2> gpurun_out/bench_v_n64.err
tail -c 600 gpurun_out/bench_v_n64.json
echo done
