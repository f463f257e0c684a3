#!/bin/bash
# fifth GPU call: full GPU test suite, bench line, reference arm, ncu launch list + full capture of the headline kernel
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_e.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_e.log
tail -3 gpurun_out/gpu_tests_e.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_e.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_e.log
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras"
timeout 600 $CMD > gpurun_out/plain_e.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_e.csv $CMD > gpurun_out/ncu_launches_e.log 2>&1
timeout 600 $CMD > gpurun_out/plain_e2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 4 -c 1 -o gpurun_out/prof_rows3_l3d_n128 $CMD > gpurun_out/ncu_full_e.log 2>&1
CMD="python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_e64.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 1 -c 1 -o gpurun_out/prof_rows3_l3d_n64 $CMD > gpurun_out/ncu_e64.log 2>&1
CMD="python tools/sweep.py --workload rmat20_n64_f64 --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_ermat.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_merge -s 1 -c 1 -o gpurun_out/prof_merge_rmat_n64 $CMD > gpurun_out/ncu_ermat.log 2>&1
CMD="python tools/sweep.py --workload band_1m_hb32_n32_f32 --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_eband.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 1 -c 1 -o gpurun_out/prof_rows3_band_n32 $CMD > gpurun_out/ncu_eband.log 2>&1
echo done
