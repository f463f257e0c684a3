#!/bin/bash
# tenth GPU call: new defaults, deep-window flavours, narrow-shape A/B, parity
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gpu_tests_j.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_j.log
tail -3 gpurun_out/gpu_tests_j.log
PV=";reg_flavour=6;reg_flavour=7;reg_flavour=5;reg_flavour=3;reg_flavour=6,rows_per_slice=32;reg_flavour=7,rows_per_slice=32;reg_flavour=6,stages=2;"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points "$PV" --out gpurun_out/sweepj_l3d_n128.jsonl > gpurun_out/sweepj_l3d_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$PV" --out gpurun_out/sweepj_l3d_n64.jsonl > gpurun_out/sweepj_l3d_n64.log 2>&1
PN=";reg_flavour=5;;reg_flavour=5;rows_per_warp=32;rows_per_warp=32,reg_flavour=5;rows_per_slice=64;rows_per_slice=64,reg_flavour=5"
for w in band_1m_hb32_n32_f32 band_1m_hb32_n1_f32 laplace2d_2048_n1_f64; do
timeout 300 python tools/sweep.py --workload $w --algo vector --steps 20 --points "$PN" --out gpurun_out/sweepj_$w.jsonl > gpurun_out/sweepj_$w.log 2>&1
done
echo done
