#!/bin/bash
# Round 2, GPU call 6 (1 GPU): the reworked substitution kernel (tests + timing) and the narrow-shape experiment
# (reg_flavour 9: two passes of rows in flight per lane group), each point in its own process.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_solve.py "tests/test_gpu_fullsize.py::test_config5_cholesky_solve_residual_check" -q -p no:cacheprovider > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c6_pytest.log
timeout 600 python - > gpurun_out/r2c6_solve.json 2> gpurun_out/r2c6_solve.err <<'PY'
import json, torch, bench
from basic_sparse_matrix_b200 import gpu
torch.cuda.set_device(0); gpu.init(0)
st = torch.cuda.Stream(); gpu.set_stream(st.cuda_stream); torch.cuda.set_stream(st)
print(json.dumps(bench.run_solve(torch, gpu)))
PY
echo "solve rc=$?"; cat gpurun_out/r2c6_solve.json; tail -3 gpurun_out/r2c6_solve.err
run() { timeout 200 python tools/sweep.py --steps 10 --algo vector --workload $1 --points "$3" --out gpurun_out/r2_pair_$1_$2.jsonl > gpurun_out/r2c6_$1_$2.log 2>&1; echo "$1 $2: $(cut -c1-100 gpurun_out/r2_pair_$1_$2.jsonl | tr '\n' ' ')"; }
for W in laplace3d_256_n4_f64 laplace3d_256_n8_f64 laplace3d_256_n16_f64; do
  run $W default ";"
  run $W pair "reg_flavour=9"
  run $W pair_s2 "reg_flavour=9,stages=2"
done
run laplace3d_256_n16_f64 pair_r32 "reg_flavour=9,rows_per_slice=32"
run band_1m_hb32_n32_f32 vec ";algo=1;algo=1,reg_flavour=9"
