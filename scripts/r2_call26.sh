#!/bin/bash
# Round 2, GPU call 26: row masks only for large results: parity of the literal call in both modes, config-1 host call time, headline e2e
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "literal or pipelined or masks or config1 or bench_as_written or kats or result_csr" > gpurun_out/r2c26_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2c26_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-target > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/r2c26_bench.json').read().strip().splitlines()[-1])
e = d['e2e']; print('e2e', e['ms_per_step'], e['parity']['bitwise'], e['d2h_bytes_per_step'], 'dense', d['e2e_dense']['ms_per_step'])
for w in d['other_workloads']:
    if 'bench_as_written' in w.get('workload', ''): print({k: v for k, v in w.items() if 'ms' in k or 'identical' in k or 'equal' in k})
PY
