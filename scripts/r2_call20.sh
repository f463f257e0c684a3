#!/bin/bash
# Round 2, GPU call 20: row-block kernel with two register tiles per lane (half as many lanes per row) on the band products
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rowblock" > gpurun_out/r2c20_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2c20_tests.log
run() { timeout 300 python tools/sweep.py --workload $1 --algo rowblock --steps 20 --points "$2" --out gpurun_out/r2_sweep_rowblock_tiles_$1.jsonl > gpurun_out/r2c20_$1.log 2>&1; echo "$1 rc=$?"; python tools/show_sweep.py gpurun_out/r2_sweep_rowblock_tiles_$1.jsonl | cut -c1-150; }
run band_1m_hb32_n64_f64 ";lanes_per_row=16;lanes_per_row=8;flags=0x7;lanes_per_row=16,flags=0x7;lanes_per_row=8,flags=0x7"
run band_1m_hb32_n128_f32 ";lanes_per_row=16;lanes_per_row=8;flags=0x7;lanes_per_row=16,flags=0x7;lanes_per_row=8,flags=0x7"
