#!/bin/bash
# row-block kernel with packed f32 multiplies (FMUL2): parity, then same-box timing
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rowblock" > gpurun_out/gpu_tests_ah.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_ah.log
tail -2 gpurun_out/gpu_tests_ah.log
timeout 600 python tools/sweep.py --workload band_1m_hb32_n32_f32 --algo auto --steps 20 --points ";algo=1;;rows_per_slice=8,algo=3" --out gpurun_out/sweepah_band_n32.jsonl > gpurun_out/sweepah_band_n32.log 2>&1
timeout 600 python tools/sweep.py --workload band_1m_hb32_n128_f32 --algo auto --steps 20 --points ";algo=1;" --out gpurun_out/sweepah_band_n128_f32.jsonl > gpurun_out/sweepah_band_n128_f32.log 2>&1
echo done
