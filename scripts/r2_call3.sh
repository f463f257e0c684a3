#!/bin/bash
# Round 2, GPU call 3 (1 GPU): the L2-prefetch experiment rebuilt WITHOUT the asm memory clobber (side library ab/libbsm_pf.so,
# selected per process with BSM_EXPERIMENT_PREFETCH=mode,rows) against the shipped library on the same box; then the e2e legs.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S="timeout 200 python tools/sweep.py --steps 10 --points ;"
for W in laplace3d_256_n128_f64 laplace3d_256_n64_f64; do
  $S --workload $W --out gpurun_out/r2_pf2_${W}_shipped.jsonl > /dev/null 2>&1
  BSM_B200_LIB=$PWD/ab/libbsm_pf.so $S --workload $W --out gpurun_out/r2_pf2_${W}_explib_off.jsonl > /dev/null 2>&1
  for M in 2,1 2,2 2,4 2,8 2,16 2,32 2,64 1,2 1,8 1,32; do
    BSM_EXPERIMENT_PREFETCH=$M BSM_B200_LIB=$PWD/ab/libbsm_pf.so $S --workload $W --out gpurun_out/r2_pf2_${W}_m${M/,/_d}.jsonl > /dev/null 2>&1
  done
done
for f in gpurun_out/r2_pf2_*.jsonl; do echo "$f $(head -1 $f | cut -c1-60)"; done
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2c3_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2c3_bench.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['phases_ms_rank0'], d['e2e'].get('pcie_probe'))
print('e2e_dense', d['e2e_dense']['ms_per_step'], d['e2e_dense']['phases_ms_rank0'])
PY
