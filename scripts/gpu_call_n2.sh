#!/bin/bash
# 2-GPU headline + x64 target on the final code (row blocks carry their global row offset for the line detection)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err; echo "exit $?" >> gpurun_out/bench_final_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e --workload laplace3d_256_n64_f64 > gpurun_out/bench_final_n64_n2.json 2> gpurun_out/bench_final_n64_n2.err; echo "exit $?" >> gpurun_out/bench_final_n64_n2.err
grep -h '^{' gpurun_out/bench_final_n2.json gpurun_out/bench_final_n64_n2.json | cut -c1-260
tail -1 gpurun_out/bench_final_n2.err
echo done
