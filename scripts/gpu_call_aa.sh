#!/bin/bash
# grouped defaults on 256/128-byte rows + flat one-tile streams (lanes_per_row == G): parity, then same-box A/B
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gpu_tests_aa.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_aa.log
tail -3 gpurun_out/gpu_tests_aa.log
P=";lanes_per_row=32;lanes_per_row=16;lanes_per_row=16,reg_flavour=5;;lanes_per_row=32;lanes_per_row=16,rows_per_slice=32;lanes_per_row=16,reg_flavour=5,rows_per_slice=32"
timeout 600 python tools/sweep.py --workload laplace3d_256_n32_f64 --algo vector --steps 10 --points "$P" --out gpurun_out/sweepaa_l3d_n32_f64.jsonl > gpurun_out/sweepaa_l3d_n32_f64.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f32 --algo vector --steps 10 --points ";lanes_per_row=32;lanes_per_row=16;" --out gpurun_out/sweepaa_l3d_n64_f32.jsonl > gpurun_out/sweepaa_l3d_n64_f32.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n16_f64 --algo vector --steps 10 --points ";lanes_per_row=32;lanes_per_row=8;lanes_per_row=8,reg_flavour=5;" --out gpurun_out/sweepaa_l3d_n16_f64.jsonl > gpurun_out/sweepaa_l3d_n16_f64.log 2>&1
PB=";lanes_per_row=8;lanes_per_row=8,reg_flavour=5;;lanes_per_row=8,rows_per_slice=32;lanes_per_row=8,reg_flavour=5,rows_per_slice=32;lanes_per_row=8,rows_per_slice=8;"
timeout 600 python tools/sweep.py --workload band_1m_hb32_n32_f32 --algo vector --steps 20 --points "$PB" --out gpurun_out/sweepaa_band_n32.jsonl > gpurun_out/sweepaa_band_n32.log 2>&1
echo done
