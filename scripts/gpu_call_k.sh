#!/bin/bash
# eleventh GPU call: 1/8 and 1/4 row blocks on one GPU (wave quantisation), new defaults everywhere, bench, ncu of final kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gpu_tests_k.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_k.log
tail -3 gpurun_out/gpu_tests_k.log
PS=";rows_per_warp=256;rows_per_warp=128;rows_per_warp=64;rows_per_warp=32"
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --slice 3/8 --algo vector --steps 20 --points "$PS" --out gpurun_out/sweepk_l3d_n128_s8.jsonl > gpurun_out/sweepk_l3d_n128_s8.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --slice 1/4 --algo vector --steps 20 --points "$PS" --out gpurun_out/sweepk_l3d_n128_s4.jsonl > gpurun_out/sweepk_l3d_n128_s4.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --slice 1/2 --algo vector --steps 20 --points "$PS" --out gpurun_out/sweepk_l3d_n128_s2.jsonl > gpurun_out/sweepk_l3d_n128_s2.log 2>&1
for w in laplace3d_256_n64_f64 band_1m_hb32_n32_f32 band_1m_hb32_n1_f32 laplace2d_2048_n1_f64 rmat20_n64_f64 rmat20_n64_f32; do
timeout 300 python tools/sweep.py --workload $w --steps 20 --points ";" --out gpurun_out/sweepk_$w.jsonl > gpurun_out/sweepk_$w.log 2>&1
done
timeout 1500 python bench.py > gpurun_out/bench_full_k.json 2> gpurun_out/bench_full_k.err; echo "bench exit $?" >> gpurun_out/bench_full_k.err
echo done
