#!/bin/bash
# round-end style run on one GPU: full GPU suite, smoke, reference arm, bench
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_final.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_final.log
tail -3 gpurun_out/gpu_tests_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_final.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref exit $?" >> gpurun_out/bench_ref_final.err
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?" >> gpurun_out/bench_final.err
du -sh gpurun_out
echo done
