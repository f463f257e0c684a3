#!/bin/bash
# Round 2, GPU call 16: ncu source-level capture of the band substitution kernels on a 2^16-row band (where does a row's time go?)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cat > gpurun_out/solve_small.py <<'PY'
import sys, os; sys.path.insert(0, os.getcwd()); import json, torch, bench
from basic_sparse_matrix_b200 import gpu
torch.cuda.set_device(0); gpu.init(0)
st = torch.cuda.Stream(); gpu.set_stream(st.cuda_stream); torch.cuda.set_stream(st)
print(json.dumps(bench.run_solve(torch, gpu, n_rows=1 << 18)))
PY
timeout 300 python gpurun_out/solve_small.py > gpurun_out/r2c16_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r2c16_plain.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trisolve_band -c 4 -o gpurun_out/r2_prof_trisolve_band -f python gpurun_out/solve_small.py > gpurun_out/r2c16_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_prof_trisolve_band.ncu-rep --page source --csv > gpurun_out/r2_prof_trisolve_band_source.csv 2>/dev/null; wc -l gpurun_out/r2_prof_trisolve_band_source.csv
ncu -i gpurun_out/r2_prof_trisolve_band.ncu-rep --page raw --csv > gpurun_out/r2_prof_trisolve_band_raw.csv 2>/dev/null
