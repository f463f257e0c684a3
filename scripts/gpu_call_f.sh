#!/bin/bash
# sixth GPU call (2 GPUs): pipelined e2e parity, bench at N=1 and N=2 (+ gather), reference arm at N=2
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_f.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or kats" > gpurun_out/gpu_tests_f.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_f.log
tail -3 gpurun_out/gpu_tests_f.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/bench_n1_f.json 2> gpurun_out/bench_n1_f.err; echo "bench1 exit $?" >> gpurun_out/bench_n1_f.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --gather > gpurun_out/bench_n2_f.json 2> gpurun_out/bench_n2_f.err; echo "bench2 exit $?" >> gpurun_out/bench_n2_f.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2_f.json 2> gpurun_out/bench_ref_n2_f.err; echo "ref2 exit $?" >> gpurun_out/bench_ref_n2_f.err
echo done
