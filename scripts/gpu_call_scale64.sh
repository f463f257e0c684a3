#!/bin/bash
# north_star target case (256^3 Laplacian x 64 columns, f64) at N GPUs
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 30 --warmup 5 --workload laplace3d_256_n64_f64 --no-e2e > gpurun_out/bench_n64_n$N.json 2> gpurun_out/bench_n64_n$N.err; echo "bench exit $?" >> gpurun_out/bench_n64_n$N.err
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --workload laplace3d_256_n64_f64 --no-e2e --no-extras --no-cpu > gpurun_out/bench_n64_n1_samebox.json 2> gpurun_out/bench_n64_n1_samebox.err
tail -2 gpurun_out/bench_n64_n$N.err | cut -c1-200
echo done
