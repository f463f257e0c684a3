#!/bin/bash
# grouped 8x4 default (16 rows x 2 stages): full GPU suite, same-box A/B of slice sizes, bench of the north_star shape, ncu capture
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_w.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_w.log
tail -3 gpurun_out/gpu_tests_w.log
P64=";lanes_per_row=32;rows_per_slice=8;rows_per_slice=32;;rows_per_slice=8,stages=3;rows_per_slice=12;rows_per_slice=16,stages=3;rows_per_slice=16,ctas_per_sm=1,warps_per_cta=20;lanes_per_row=32;"
timeout 900 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$P64" --out gpurun_out/sweepw_l3d_n64.jsonl > gpurun_out/sweepw_l3d_n64.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --slice 3/8 --points ";lanes_per_row=32;;rows_per_warp=256;rows_per_warp=64" --out gpurun_out/sweepw_l3d_n64_s8.jsonl > gpurun_out/sweepw_l3d_n64_s8.log 2>&1
timeout 900 python bench.py --workload laplace3d_256_n64_f64 --steps 30 --warmup 5 --no-extras --no-cpu > gpurun_out/bench_w_n64.json 2> gpurun_out/bench_w_n64.err
tail -c 300 gpurun_out/bench_w_n64.json
timeout 300 python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 1 --warmup 1 > gpurun_out/plain_w.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:spmm_rows -s 1 -c 1 -o /tmp/prof_w python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 1 --warmup 1 > gpurun_out/ncu_w.log 2>&1
ncu -i /tmp/prof_w.ncu-rep --page raw --csv > gpurun_out/prof_grouped_l3d_n64.raw.csv 2>/dev/null
du -sh gpurun_out
echo done
