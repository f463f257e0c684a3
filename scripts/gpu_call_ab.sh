#!/bin/bash
# A/B of two builds of the library on one box: ab/libbsm_prev.so vs the in-tree build
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gpu_tests_ab.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_ab.log
tail -3 gpurun_out/gpu_tests_ab.log
for w in laplace3d_256_n128_f64 laplace3d_256_n64_f64 laplace2d_2048_n1_f64 band_1m_hb32_n32_f32 band_1m_hb32_n1_f32 rmat20_n64_f64; do
BSM_B200_LIB=$PWD/ab/libbsm_prev.so timeout 300 python tools/sweep.py --workload $w --steps 10 --points ";" --out gpurun_out/ab_prev_${w}.jsonl > gpurun_out/ab_prev_${w}.log 2>&1
timeout 300 python tools/sweep.py --workload $w --steps 10 --points ";" --out gpurun_out/ab_new_${w}.jsonl > gpurun_out/ab_new_${w}.log 2>&1
done
echo done
