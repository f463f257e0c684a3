#!/bin/bash
# scaling call: bench at N = number of visible GPUs (+ gather), reference arm rank-0 only
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --gather > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?" >> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err | cut -c1-200
echo done
