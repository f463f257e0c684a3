#!/bin/bash
# Round 2, GPU call 27: band substitution kernels with 16 lanes per column for hb 32 (second build) against 8 (shipped), same box
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cat > gpurun_out/solve_full.py <<'PY'
import sys, os; sys.path.insert(0, os.getcwd()); import json, torch, bench
from basic_sparse_matrix_b200 import gpu
torch.cuda.set_device(0); gpu.init(0)
st = torch.cuda.Stream(); gpu.set_stream(st.cuda_stream); torch.cuda.set_stream(st)
r = bench.run_solve(torch, gpu, n_rows=1 << 19)
print(json.dumps({k: r[k] for k in ("forward_ms", "backward_ms", "x_equals_cpu_port_bitwise")}))
PY
for v in lib_exp lib_exp2 lib_exp lib_exp2; do
  BSM_B200_LIB=$PWD/basic_sparse_matrix_b200/$v/libbsm_b200.so timeout 300 python gpurun_out/solve_full.py 2>&1 | tail -1 | sed "s/^/$v: /"
done
