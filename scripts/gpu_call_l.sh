#!/bin/bash
# twelfth GPU call: pool allocator validation (full GPU suite), bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_l.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_l.log
tail -3 gpurun_out/gpu_tests_l.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_l.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_l.log
timeout 1500 python bench.py > gpurun_out/bench_full_l.json 2> gpurun_out/bench_full_l.err; echo "bench exit $?" >> gpurun_out/bench_full_l.err
PS=";rows_per_warp=256;rows_per_warp=128"
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --slice 3/8 --algo vector --steps 20 --points "$PS" --out gpurun_out/sweepl_l3d_n128_s8.jsonl > gpurun_out/sweepl_l3d_n128_s8.log 2>&1
echo done
