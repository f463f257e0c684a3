#!/bin/bash
# thirteenth GPU call: merge-path explorations + ncu of the final merge / rows kernels
set -u
mkdir -p gpurun_out
PM=";col_tile=32,prefer_wide_rows=-1;col_tile=16,prefer_wide_rows=-1;prefer_wide_rows=-1;warps_per_cta=4;warps_per_cta=2;flags=0x80000000;flags=0x80000002"
timeout 600 python tools/sweep.py --workload rmat20_n64_f64 --algo merge --steps 8 --points "$PM" --out gpurun_out/sweepm_rmat_f64.jsonl > gpurun_out/sweepm_rmat_f64.log 2>&1
timeout 600 python tools/sweep.py --workload rmat20_n64_f32 --algo merge --steps 8 --points "$PM" --out gpurun_out/sweepm_rmat_f32.jsonl > gpurun_out/sweepm_rmat_f32.log 2>&1
CMD="python tools/sweep.py --workload rmat20_n64_f64 --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_mrmat.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_merge -s 1 -c 1 -o gpurun_out/prof_merge2_rmat_n64 $CMD > gpurun_out/ncu_mrmat.log 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras"
timeout 600 $CMD > gpurun_out/plain_m.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_m.csv $CMD > gpurun_out/ncu_launches_m.log 2>&1
timeout 600 $CMD > gpurun_out/plain_m2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 4 -c 1 -o gpurun_out/prof_rows4_l3d_n128 $CMD > gpurun_out/ncu_full_m.log 2>&1
CMD="python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_m64.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 1 -c 1 -o gpurun_out/prof_rows4_l3d_n64 $CMD > gpurun_out/ncu_m64.log 2>&1
echo done
