#!/bin/bash
# interleaved rows between lane groups (interleave_rows=1): parity, then same-box A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped or near_diagonal or tuning" > gpurun_out/gpu_tests_ilv.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_ilv.log
tail -3 gpurun_out/gpu_tests_ilv.log
P64=";lanes_per_row=8,reg_flavour=7,interleave_rows=1,rows_per_slice=16,stages=2;lanes_per_row=8,reg_flavour=7,interleave_rows=1,rows_per_slice=32,stages=2;lanes_per_row=8,reg_flavour=5,interleave_rows=1,rows_per_slice=16;;lanes_per_row=8,reg_flavour=7,interleave_rows=1,rows_per_slice=8,stages=2;lanes_per_row=16,reg_flavour=7,interleave_rows=1,rows_per_slice=16,stages=2;lanes_per_row=8,reg_flavour=7,interleave_rows=1,rows_per_slice=16,stages=3"
timeout 900 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 10 --points "$P64" --out gpurun_out/sweepilv_l3d_n64.jsonl > gpurun_out/sweepilv_l3d_n64.log 2>&1
P128=";lanes_per_row=16,reg_flavour=7,interleave_rows=1,rows_per_slice=16,stages=2;lanes_per_row=16,reg_flavour=5,interleave_rows=1;lanes_per_row=16,reg_flavour=7,interleave_rows=1,rows_per_slice=8,stages=2;;lanes_per_row=16,reg_flavour=7,interleave_rows=1,rows_per_slice=32,stages=2"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points "$P128" --out gpurun_out/sweepilv_l3d_n128.jsonl > gpurun_out/sweepilv_l3d_n128.log 2>&1
P32=";lanes_per_row=8,reg_flavour=7,interleave_rows=1,rows_per_slice=32,stages=2;lanes_per_row=8,reg_flavour=7,interleave_rows=1,rows_per_slice=16,stages=2;lanes_per_row=4,reg_flavour=7,interleave_rows=1,rows_per_slice=32,stages=2;"
timeout 600 python tools/sweep.py --workload laplace3d_256_n32_f64 --algo vector --steps 10 --points "$P32" --out gpurun_out/sweepilv_l3d_n32_f64.jsonl > gpurun_out/sweepilv_l3d_n32_f64.log 2>&1
PB=";lanes_per_row=4,interleave_rows=1;lanes_per_row=4,interleave_rows=1,reg_flavour=7;lanes_per_row=4,interleave_rows=1,rows_per_slice=8;;lanes_per_row=4,interleave_rows=1,rows_per_slice=32"
timeout 600 python tools/sweep.py --workload band_1m_hb32_n32_f32 --algo vector --steps 20 --points "$PB" --out gpurun_out/sweepilv_band_n32.jsonl > gpurun_out/sweepilv_band_n32.log 2>&1
echo done
