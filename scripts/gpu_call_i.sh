#!/bin/bash
# ninth GPU call: full GPU tests (config 5 residual check), A/B of vectorised vs scalar A-stream reads, bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_i.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_i.log
tail -3 gpurun_out/gpu_tests_i.log
PV=";reg_flavour=5;reg_flavour=3;reg_flavour=5;rows_per_slice=32;rows_per_slice=32,reg_flavour=5;stages=2;stages=2,reg_flavour=5;reg_flavour=3;reg_flavour=5"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points "$PV" --out gpurun_out/sweepi_l3d_n128.jsonl > gpurun_out/sweepi_l3d_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$PV" --out gpurun_out/sweepi_l3d_n64.jsonl > gpurun_out/sweepi_l3d_n64.log 2>&1
timeout 1500 python bench.py > gpurun_out/bench_full_i.json 2> gpurun_out/bench_full_i.err; echo "bench exit $?" >> gpurun_out/bench_full_i.err
echo done
