#!/bin/bash
# parity + narrow-shape sweeps after the cursor bookkeeping / predicated prologue changes; bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_p.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_p.log
tail -3 gpurun_out/gpu_tests_p.log
PN=";rows_per_slice=32;rows_per_slice=64;rows_per_slice=128;rows_per_slice=32,stages=3;rows_per_slice=32,stages=4;rows_per_warp=256,rows_per_slice=32;warps_per_cta=8,ctas_per_sm=2"
for w in laplace2d_2048_n1_f64 band_1m_hb32_n1_f32 band_1m_hb32_n32_f32; do
timeout 300 python tools/sweep.py --workload $w --algo vector --steps 20 --points "$PN" --out gpurun_out/sweepp_$w.jsonl > gpurun_out/sweepp_$w.log 2>&1
done
for w in laplace3d_256_n128_f64 laplace3d_256_n64_f64; do
timeout 300 python tools/sweep.py --workload $w --steps 10 --points ";" --out gpurun_out/sweepp_$w.jsonl > gpurun_out/sweepp_$w.log 2>&1
done
echo done
