#!/bin/bash
# full GPU suite + smoke + default bench on the final state of the round; ncu capture of the row-block kernel
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_ai.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_ai.log
tail -3 gpurun_out/gpu_tests_ai.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_ai.log 2>&1; tail -1 gpurun_out/smoke_ai.log
timeout 1200 python bench.py > gpurun_out/bench_ai.json 2> gpurun_out/bench_ai.err
tail -c 200 gpurun_out/bench_ai.json
timeout 300 python tools/sweep.py --workload band_1m_hb32_n32_f32 --steps 1 --warmup 1 > gpurun_out/plain_ai.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:spmm_rowblock -s 1 -c 1 -o /tmp/prof_ai python tools/sweep.py --workload band_1m_hb32_n32_f32 --steps 1 --warmup 1 > gpurun_out/ncu_ai.log 2>&1
ncu -i /tmp/prof_ai.ncu-rep --page raw --csv > gpurun_out/prof_rowblock_band_n32.raw.csv 2>/dev/null
echo done
