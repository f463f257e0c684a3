#!/bin/bash
# Round 2, GPU call 19: host-side column fill for full result blocks of the literal call: parity tests, then the e2e lines of bench.py with / without
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo skip tests
nproc; free -g | head -2
for cfg in "8 1" "4 1" "6 1" "12 1"; do
  set -- $cfg; t=$1; nt=$2
  BSM_PIPE_EXPAND_THREADS=$t BSM_PIPE_EXPAND_NT=$nt timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-target > gpurun_out/r2c19_bench_t${t}_nt$nt.json 2> gpurun_out/r2c19_bench_t${t}_nt$nt.err; echo "bench t=$t nt=$nt rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2c19_bench_t${t}_nt$nt.json').read().strip().splitlines()[-1])
e = d['e2e']
print('threads $t nt $nt: e2e', e['ms_per_step'], 'ms', e['value'], 'GFLOP/s d2h', e['d2h_bytes_per_step'], 'parity', e['parity']['bitwise'], {k: v for k, v in e.get('phases_ms_rank0', {}).items() if k in ('b_h2d_device','d2h_device','wait_host','total_host')}, 'dense', d['e2e_dense']['ms_per_step'])
PY
done
