#!/bin/bash
# Round 2, GPU call 7 (1 GPU): the row-pipelined substitution kernel (tests + timing), the statistics test.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_solve.py "tests/test_gpu_fullsize.py::test_config5_cholesky_solve_residual_check" "tests/test_gpu_parity.py::test_line_length_is_a_majority_vote_over_all_rows" -q -p no:cacheprovider > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c7_pytest.log
timeout 600 python - > gpurun_out/r2c7_solve.json 2> gpurun_out/r2c7_solve.err <<'PY'
import json, torch, bench
from basic_sparse_matrix_b200 import gpu
torch.cuda.set_device(0); gpu.init(0)
st = torch.cuda.Stream(); gpu.set_stream(st.cuda_stream); torch.cuda.set_stream(st)
print(json.dumps(bench.run_solve(torch, gpu)))
print(json.dumps(bench.run_solve(torch, gpu, nrhs=128)))
PY
echo "solve rc=$?"; cat gpurun_out/r2c7_solve.json; tail -3 gpurun_out/r2c7_solve.err
