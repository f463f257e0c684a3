#!/bin/bash
# after the spmm_vector refactor: full GPU suite, smoke, default bench (refreshes profiles/r1_bench_n1.json), reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_ac.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_ac.log
tail -3 gpurun_out/gpu_tests_ac.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_ac.log 2>&1; tail -2 gpurun_out/smoke_ac.log
timeout 1200 python bench.py > gpurun_out/bench_ac.json 2> gpurun_out/bench_ac.err
tail -c 200 gpurun_out/bench_ac.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ac_reference.json 2> gpurun_out/bench_ac_reference.err
tail -c 300 gpurun_out/bench_ac_reference.json
echo done
