#!/bin/bash
# 2-lane grouped shapes (64-byte output rows): parity, then same-box A/B on the stencil x8 f64 / x16 f32
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped or random_shape or bitwise_random" > gpurun_out/gpu_tests_ak.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_ak.log
tail -2 gpurun_out/gpu_tests_ak.log
timeout 600 python tools/sweep.py --workload laplace3d_256_n8_f64 --algo vector --steps 10 --points ";lanes_per_row=32;;lanes_per_row=32;lanes_per_row=2,reg_flavour=7;lanes_per_row=2,rows_per_slice=64" --out gpurun_out/sweepak_l3d_n8_f64.jsonl > gpurun_out/sweepak_l3d_n8_f64.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_ak.log 2>&1; tail -1 gpurun_out/smoke_ak.log
echo done
