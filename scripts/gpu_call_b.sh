#!/bin/bash
# second GPU call: parity of the rewritten vector kernel, memory-system probe, tuning sweeps
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gpu_tests_b.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_b.log
tail -3 gpurun_out/gpu_tests_b.log
timeout 300 ./tools/membw > gpurun_out/membw.log 2>&1; echo "membw exit $?" >> gpurun_out/membw.log
P128=";sync_rows=-1;rows_per_warp=16,sync_rows=-1;col_tile=64;col_tile=64,sync_rows=-1;col_tile=64,rows_per_warp=16,sync_rows=-1;col_tile=32;col_tile=32,sync_rows=-1;col_tile=16;col_tile=64,sync_rows=2;col_tile=64,sync_rows=8;col_tile=64,sync_rows=16;col_tile=64,warps_per_cta=8;col_tile=64,warps_per_cta=8,ctas_per_sm=1;col_tile=64,rows_per_slice=8;col_tile=64,rows_per_slice=32;col_tile=64,stages=2;col_tile=64,stages=4;col_tile=64,flags=0x80000000;col_tile=64,flags=0x80000001;col_tile=64,flags=0x80000002;col_tile=64,rows_per_warp=128;col_tile=64,rows_per_warp=512;col_tile=64,rows_per_warp=1024,sync_rows=-1;col_tile=64,rows_per_warp=64,sync_rows=-1"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 8 --points "$P128" --out gpurun_out/sweep_l3d_n128.jsonl > gpurun_out/sweep_l3d_n128.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep_l3d_n128.log
P64=";sync_rows=-1;rows_per_warp=16,sync_rows=-1;col_tile=32;col_tile=32,sync_rows=-1;sync_rows=2;sync_rows=8;warps_per_cta=8;rows_per_slice=8;rows_per_slice=32"
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 8 --points "$P64" --out gpurun_out/sweep_l3d_n64.jsonl > gpurun_out/sweep_l3d_n64.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep_l3d_n64.log
PS=";sync_rows=-1;rows_per_warp=32,sync_rows=-1;rows_per_slice=64;rows_per_slice=128;warps_per_cta=8;warps_per_cta=4,ctas_per_sm=4;prefer_wide_rows=1"
timeout 300 python tools/sweep.py --workload laplace2d_2048_n1_f64 --algo vector --steps 20 --points "$PS" --out gpurun_out/sweep_l2d_n1.jsonl > gpurun_out/sweep_l2d_n1.log 2>&1
timeout 300 python tools/sweep.py --workload band_1m_hb32_n32_f32 --algo vector --steps 20 --points "$PS" --out gpurun_out/sweep_band_n32.jsonl > gpurun_out/sweep_band_n32.log 2>&1
PM=";merge_items=128;merge_items=512;merge_items=1024;col_tile=32;warps_per_cta=4"
timeout 600 python tools/sweep.py --workload rmat20_n64_f64 --algo merge --steps 8 --points "$PM" --out gpurun_out/sweep_rmat_f64.jsonl > gpurun_out/sweep_rmat_f64.log 2>&1
echo done
