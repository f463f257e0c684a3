#!/bin/bash
# third GPU call: parity of the rolling-window vector kernel, flavour sweeps, ncu captures
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gpu_tests_c.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_c.log
tail -3 gpurun_out/gpu_tests_c.log
P128=";reg_flavour=2;reg_flavour=3;reg_flavour=4;reg_flavour=3,rows_per_warp=16;reg_flavour=4,rows_per_warp=16;rows_per_warp=16;col_tile=64;col_tile=64,reg_flavour=2;col_tile=64,reg_flavour=3;col_tile=64,reg_flavour=4;rows_per_slice=32;rows_per_slice=64;reg_flavour=3,rows_per_slice=32;reg_flavour=4,rows_per_slice=32;warps_per_cta=8;warps_per_cta=12;rows_per_warp=128;rows_per_warp=512;reg_flavour=4,rows_per_warp=128;reg_flavour=4,rows_per_warp=512;reg_flavour=4,stages=2;reg_flavour=4,flags=0x80000000;reg_flavour=4,flags=0x80000001;reg_flavour=4,warps_per_cta=4,ctas_per_sm=8"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 8 --points "$P128" --out gpurun_out/sweepc_l3d_n128.jsonl > gpurun_out/sweepc_l3d_n128.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweepc_l3d_n128.log
P64=";reg_flavour=2;reg_flavour=3;reg_flavour=4;reg_flavour=4,rows_per_warp=16;rows_per_warp=16;rows_per_slice=32;reg_flavour=4,rows_per_slice=32;reg_flavour=4,rows_per_slice=64;warps_per_cta=8;reg_flavour=4,rows_per_warp=128;reg_flavour=4,rows_per_warp=512;reg_flavour=4,warps_per_cta=4,ctas_per_sm=8;reg_flavour=3,warps_per_cta=4,ctas_per_sm=6"
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 8 --points "$P64" --out gpurun_out/sweepc_l3d_n64.jsonl > gpurun_out/sweepc_l3d_n64.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweepc_l3d_n64.log
PS=";rows_per_slice=64;rows_per_slice=128;rows_per_slice=256;warps_per_cta=8;rows_per_warp=2048;rows_per_warp=256;rows_per_warp=256,rows_per_slice=64"
timeout 300 python tools/sweep.py --workload laplace2d_2048_n1_f64 --algo vector --steps 20 --points "$PS" --out gpurun_out/sweepc_l2d_n1.jsonl > gpurun_out/sweepc_l2d_n1.log 2>&1
PB=";rows_per_warp=32;rows_per_warp=64;rows_per_warp=128;rows_per_warp=64,rows_per_slice=8;rows_per_warp=64,warps_per_cta=8;prefer_wide_rows=1,rows_per_warp=64"
timeout 300 python tools/sweep.py --workload band_1m_hb32_n32_f32 --algo vector --steps 20 --points "$PB" --out gpurun_out/sweepc_band_n32.jsonl > gpurun_out/sweepc_band_n32.log 2>&1
timeout 300 python tools/sweep.py --workload band_1m_hb32_n1_f32 --algo vector --steps 20 --points "$PB" --out gpurun_out/sweepc_band_n1.jsonl > gpurun_out/sweepc_band_n1.log 2>&1
CMD="python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_c64.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 1 -c 1 -o gpurun_out/prof_rows2_l3d_n64 $CMD > gpurun_out/ncu_c64.log 2>&1
CMD="python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_c128.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 1 -c 1 -o gpurun_out/prof_rows2_l3d_n128 $CMD > gpurun_out/ncu_c128.log 2>&1
echo done
