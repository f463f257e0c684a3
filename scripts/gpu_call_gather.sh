#!/bin/bash
# N-GPU gathered-product comparison: bsm_spmm + NCCL all-gather vs the fused scatter kernel
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --gather --no-e2e > gpurun_out/bench_gather_n$N.json 2> gpurun_out/bench_gather_n$N.err; echo "bench exit $?" >> gpurun_out/bench_gather_n$N.err
tail -2 gpurun_out/bench_gather_n$N.err | cut -c1-300
echo done
