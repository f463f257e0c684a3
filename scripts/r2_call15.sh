#!/bin/bash
# Round 2, GPU call 15: band substitution kernels, quick loop: band-kernel parity tests + config 5 timing
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_solve.py -x -q -m gpu -k "band" > gpurun_out/r2c15_solve_tests.log 2>&1; echo "band tests rc=$?"; tail -3 gpurun_out/r2c15_solve_tests.log
cat > gpurun_out/solve_full.py <<'PY'
import sys, os; sys.path.insert(0, os.getcwd()); import json, torch, bench
from basic_sparse_matrix_b200 import gpu
torch.cuda.set_device(0); gpu.init(0)
st = torch.cuda.Stream(); gpu.set_stream(st.cuda_stream); torch.cuda.set_stream(st)
print(json.dumps(bench.run_solve(torch, gpu)))
PY
timeout 600 python gpurun_out/solve_full.py > gpurun_out/r2c15_solve_band.log 2>&1; echo "band rc=$?"; tail -1 gpurun_out/r2c15_solve_band.log | cut -c1-400
