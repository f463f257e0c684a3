#!/bin/bash
# Round 2, GPU call 13: flat narrow streams as a kernel variant of their own (reg_flavour 9): parity, then the default on x8 f64.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu  > gpurun_out/r2c13_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2c13_parity.log
timeout 300 python tools/sweep.py --workload laplace3d_256_n8_f64 --steps 10 --points ";reg_flavour=1;reg_flavour=9,rows_per_warp=128;reg_flavour=9,rows_per_slice=32;" --out gpurun_out/r2_sweep_flat_narrow_l3d_n8.jsonl > gpurun_out/r2c13_n8.log 2>&1; echo "n8 rc=$?"
python tools/show_sweep.py gpurun_out/r2_sweep_flat_narrow_l3d_n8.jsonl | cut -c1-175
