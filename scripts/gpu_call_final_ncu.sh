#!/bin/bash
# ncu evidence for the final kernels: full captures converted to raw CSV on the box (reports are too big to ship)
set -u
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 300 "$@" > gpurun_out/plain_$name.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none -k regex:$re -s $skip -c 1 -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_$name.raw.csv 2>/dev/null
}
cap final_l3d_n64 spmm_rows 1 python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 1 --warmup 1
cap final_l2d_n1 spmm_rows 1 python tools/sweep.py --workload laplace2d_2048_n1_f64 --steps 1 --warmup 1
cap final_rmat_f64 spmm_merge 1 python tools/sweep.py --workload rmat20_n64_f64 --steps 1 --warmup 1
du -sh gpurun_out
echo done
