#!/bin/bash
# Round 2, GPU call 2 (1 GPU): same-box A/B of the round-1 library against the current one on the shapes that moved between
# the rounds, the GPU test-suite, a full default bench run, the ncu launch list of that bench and full ncu captures of the
# dominant kernels (each only after the plain run of the same command exited 0).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S="timeout 300 python tools/sweep.py --steps 10"
for W in laplace3d_256_n128_f64 laplace2d_2048_n1_f64 laplace3d_252_n128_f64 laplace3d_256_n64_f64; do
  BSM_B200_LIB=$PWD/ab/libbsm_r1.so $S --workload $W --points ";" --out gpurun_out/r2_ab_r1lib_$W.jsonl > gpurun_out/r2c2_ab1_$W.log 2>&1
  $S --workload $W --points ";" --out gpurun_out/r2_ab_r2lib_$W.jsonl > gpurun_out/r2c2_ab2_$W.log 2>&1
  BSM_B200_LIB=$PWD/ab/libbsm_r1.so $S --workload $W --points ";" --out gpurun_out/r2_ab_r1lib_again_$W.jsonl > gpurun_out/r2c2_ab3_$W.log 2>&1
done
for f in gpurun_out/r2_ab_*.jsonl; do echo $f; cut -c1-90 $f; done
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log )
tail -4 gpurun_out/r2c2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench rc=$?"
tail -c 400 gpurun_out/r2c2_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2c2_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'parity', 'north_star_target')})
print('e2e', d['e2e'])
print('e2e_dense', d['e2e_dense'])
PY
# ncu: launch list of a short bench run (kernel shares), then full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2c2_ncu_launch.log 2>&1
echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_kernel -s 4 -c 1 -o gpurun_out/r2_prof_rows_l3d_n128 -f python tools/sweep.py --workload laplace3d_256_n128_f64 --steps 2 --warmup 2 --points ";" > gpurun_out/r2c2_ncu1.log 2>&1
echo "ncu rows n128 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_kernel -s 4 -c 1 -o gpurun_out/r2_prof_rows_l3d_n64 -f python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 2 --warmup 2 --points ";" > gpurun_out/r2c2_ncu2.log 2>&1
echo "ncu rows n64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_merge_kernel -s 4 -c 1 -o gpurun_out/r2_prof_merge_rmat_f64 -f python tools/sweep.py --workload rmat20_n64_f64 --steps 2 --warmup 2 --points ";" > gpurun_out/r2c2_ncu3.log 2>&1
echo "ncu merge rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rowblock_kernel -s 4 -c 1 -o gpurun_out/r2_prof_rowblock_fused_band_n32 -f python tools/sweep.py --workload band_1m_hb32_n32_f32 --steps 2 --warmup 2 --points "flags=7" > gpurun_out/r2c2_ncu4.log 2>&1
echo "ncu rowblock fused rc=$?"
ls -la gpurun_out/*.ncu-rep
