#!/bin/bash
# 2-GPU call: scatter parity on one GPU, then bench N=2 with the NCCL gather and the fused scatter
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scatter or kats or tuning" > gpurun_out/gpu_tests_o.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_o.log
tail -4 gpurun_out/gpu_tests_o.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --gather --no-e2e > gpurun_out/bench_n2_o.json 2> gpurun_out/bench_n2_o.err; echo "bench exit $?" >> gpurun_out/bench_n2_o.err
tail -3 gpurun_out/bench_n2_o.err | cut -c1-300
echo done
