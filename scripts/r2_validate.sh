#!/bin/bash
# Round 2: what the driver runs at round end, on one GPU: the GPU test-suite, smoke(), the default bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log )
tail -4 gpurun_out/r2v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2v_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2v_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['parity']['bitwise'], d['roofline']['frac'], d['roofline']['traffic'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['parity']['bitwise'], 'dense', d['e2e_dense']['ms_per_step'])
for w in d['other_workloads']:
    if 'solve' in w.get('workload', '') or 'error' in w: print(w)
PY
