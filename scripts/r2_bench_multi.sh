#!/bin/bash
# Round 2: bench.py on N GPUs of one box exactly as the driver launches it (torchrun, one rank per GPU), incl. the
# gathered result with oracle parity, the north_star target case and both e2e forms with per-phase timers.
#   bash scripts/r2_bench_multi.sh N [extra bench.py args]
set -u
N=${1:-2}
shift || true
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/r2_topo_n${N}.txt 2>&1
numactl -H > gpurun_out/r2_numa_n${N}.txt 2>&1 || lscpu | grep -i numa > gpurun_out/r2_numa_n${N}.txt 2>&1
free -g >> gpurun_out/r2_numa_n${N}.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err
echo "bench N=$N rc=$?"
tail -c 1500 gpurun_out/r2_bench_n${N}.err
head -c 6000 gpurun_out/r2_bench_n${N}.json
