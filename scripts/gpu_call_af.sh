#!/bin/bash
# row-block kernel for band-like matrices: parity, then same-box A/B on the config-5 band
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rowblock or kats or scatter" > gpurun_out/gpu_tests_af.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_af.log
tail -3 gpurun_out/gpu_tests_af.log
timeout 600 python tools/sweep.py --workload band_1m_hb32_n32_f32 --algo auto --steps 20 --points ";algo=1;;algo=1;algo=3,col_tile=16" --out gpurun_out/sweepaf_band_n32.jsonl > gpurun_out/sweepaf_band_n32.log 2>&1
timeout 600 python tools/sweep.py --workload band_1m_hb32_n1_f32 --algo auto --steps 20 --points ";algo=1;algo=3;;algo=1;algo=3" --out gpurun_out/sweepaf_band_n1.jsonl > gpurun_out/sweepaf_band_n1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config5" > gpurun_out/gpu_tests_af_full.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_af_full.log
tail -3 gpurun_out/gpu_tests_af_full.log
echo done
