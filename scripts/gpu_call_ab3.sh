#!/bin/bash
# A/B on one box, interleaved x3: B-row loads with L2 evict_last hint vs default
set -u
mkdir -p gpurun_out
for rep in 1 2 3; do
for w in laplace3d_256_n128_f64 laplace3d_256_n64_f64 rmat20_n64_f64; do
BSM_B200_LIB=$PWD/ab/libbsm_evict_last.so timeout 300 python tools/sweep.py --workload $w --steps 10 --points "" --out gpurun_out/ab3_evl_${w}_$rep.jsonl > gpurun_out/ab3_evl_${w}_$rep.log 2>&1
timeout 300 python tools/sweep.py --workload $w --steps 10 --points "" --out gpurun_out/ab3_def_${w}_$rep.jsonl > gpurun_out/ab3_def_${w}_$rep.log 2>&1
done
done
echo done
