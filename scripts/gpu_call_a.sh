#!/bin/bash
# first GPU call: parity tests, smoke, full bench line, ncu launch list + one full capture of the top kernel
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 4 -c 1 -o gpurun_out/prof_rows_l3d_n128 $CMD > gpurun_out/ncu_full.log 2>&1
echo done
