#!/bin/bash
# Round 2, GPU call 14: band substitution kernels (one solver warp + staging warps): parity, then config 5 timing against the general kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_solve.py -x -q -m gpu > gpurun_out/r2c14_solve_tests.log 2>&1; echo "solve tests rc=$?"; tail -15 gpurun_out/r2c14_solve_tests.log
cat > gpurun_out/solve_full.py <<'PY'
import sys, os; sys.path.insert(0, os.getcwd()); import json, torch, bench
from basic_sparse_matrix_b200 import gpu
torch.cuda.set_device(0); gpu.init(0)
st = torch.cuda.Stream(); gpu.set_stream(st.cuda_stream); torch.cuda.set_stream(st)
print(json.dumps(bench.run_solve(torch, gpu)))
PY
timeout 600 python gpurun_out/solve_full.py > gpurun_out/r2c14_solve_band.log 2>&1; echo "band rc=$?"; tail -1 gpurun_out/r2c14_solve_band.log | cut -c1-1500
BSM_SOLVE_GENERAL=1 timeout 600 python gpurun_out/solve_full.py > gpurun_out/r2c14_solve_general.log 2>&1; echo "general rc=$?"; tail -1 gpurun_out/r2c14_solve_general.log | cut -c1-600
