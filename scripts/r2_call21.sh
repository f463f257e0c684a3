#!/bin/bash
# Round 2, GPU call 21: row-block kernel, two register tiles per lane as the default: parity of everything that runs it, band workloads
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_build_flags.py -x -q -m "gpu or not gpu" -k "rowblock or band or config5 or build or edge or random_shape or literal or pipelined" > gpurun_out/r2c21_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c21_tests.log
run() { timeout 300 python tools/sweep.py --workload $1 --algo auto --steps 20 --points "$2" --out gpurun_out/r2_sweep_rowblock_default_$1.jsonl > gpurun_out/r2c21_$1.log 2>&1; echo "$1 rc=$?"; python tools/show_sweep.py gpurun_out/r2_sweep_rowblock_default_$1.jsonl | cut -c1-150; }



