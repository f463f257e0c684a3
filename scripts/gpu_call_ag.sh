#!/bin/bash
# row-block kernel: 4 vs 8 rows per lane group, on the config-5 band (x32 f32) and on wider / f64 bands
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rowblock" > gpurun_out/gpu_tests_ag.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_ag.log
tail -2 gpurun_out/gpu_tests_ag.log
timeout 600 python tools/sweep.py --workload band_1m_hb32_n32_f32 --algo auto --steps 20 --points ";rows_per_slice=4,algo=3;rows_per_slice=8,algo=3;algo=1;" --out gpurun_out/sweepag_band_n32.jsonl > gpurun_out/sweepag_band_n32.log 2>&1
timeout 600 python tools/sweep.py --workload band_1m_hb32_n64_f64 --algo auto --steps 20 --points ";rows_per_slice=4,algo=3;rows_per_slice=8,algo=3;algo=1;" --out gpurun_out/sweepag_band_n64_f64.jsonl > gpurun_out/sweepag_band_n64_f64.log 2>&1
timeout 600 python tools/sweep.py --workload band_1m_hb32_n128_f32 --algo auto --steps 20 --points ";rows_per_slice=4,algo=3;rows_per_slice=8,algo=3;algo=1;" --out gpurun_out/sweepag_band_n128_f32.jsonl > gpurun_out/sweepag_band_n128_f32.log 2>&1
echo done
