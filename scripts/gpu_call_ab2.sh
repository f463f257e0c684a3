#!/bin/bash
# A/B on one box, interleaved and repeated: merge kernel LDS.128 vs scalar A reads; n=64 default (24-warp CTA) vs 3x8
set -u
mkdir -p gpurun_out
for rep in 1 2 3; do
for w in rmat20_n64_f64 rmat20_n64_f32; do
BSM_B200_LIB=$PWD/ab/libbsm_merge_scalar.so timeout 300 python tools/sweep.py --workload $w --steps 10 --points "" --out gpurun_out/ab2_scalar_${w}_$rep.jsonl > gpurun_out/ab2_scalar_${w}_$rep.log 2>&1
timeout 300 python tools/sweep.py --workload $w --steps 10 --points "" --out gpurun_out/ab2_vec_${w}_$rep.jsonl > gpurun_out/ab2_vec_${w}_$rep.log 2>&1
done
done
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 10 --points ";reg_flavour=3;;reg_flavour=3;;reg_flavour=5;;reg_flavour=5" --out gpurun_out/ab2_n64.jsonl > gpurun_out/ab2_n64.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --steps 10 --points ";reg_flavour=7;;reg_flavour=7;;reg_flavour=7" --out gpurun_out/ab2_n128.jsonl > gpurun_out/ab2_n128.log 2>&1
echo done
