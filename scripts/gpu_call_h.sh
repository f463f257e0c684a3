#!/bin/bash
# eighth GPU call: parity after the fix-up rewrite, slice/stage sweeps, merge-path sweeps
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_h.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_h.log
tail -3 gpurun_out/gpu_tests_h.log
PV=";stages=2;rows_per_slice=24;rows_per_slice=24,stages=2;rows_per_slice=32;rows_per_slice=32,stages=2;rows_per_slice=48,stages=2;rows_per_slice=64,stages=2;rows_per_slice=32,stages=2,rows_per_warp=128;stages=4"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 8 --points "$PV" --out gpurun_out/sweeph_l3d_n128.jsonl > gpurun_out/sweeph_l3d_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 8 --points "$PV" --out gpurun_out/sweeph_l3d_n64.jsonl > gpurun_out/sweeph_l3d_n64.log 2>&1
PM=";merge_items=320;merge_items=384;merge_items=448;merge_items=512;merge_items=640;warps_per_cta=4,merge_items=384;warps_per_cta=4,merge_items=768"
timeout 600 python tools/sweep.py --workload rmat20_n64_f64 --algo merge --steps 8 --points "$PM" --out gpurun_out/sweeph_rmat_f64.jsonl > gpurun_out/sweeph_rmat_f64.log 2>&1
timeout 600 python tools/sweep.py --workload rmat20_n64_f32 --algo merge --steps 8 --points "$PM" --out gpurun_out/sweeph_rmat_f32.jsonl > gpurun_out/sweeph_rmat_f32.log 2>&1
CMD="python tools/sweep.py --workload rmat20_n64_f64 --steps 2 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_hrmat.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_h_rmat.csv $CMD > gpurun_out/ncu_launches_h.log 2>&1
echo done
