#!/bin/bash
# Round 2, GPU call 23: ncu captures of the kernels that changed in the second session (after their plain runs exited 0 earlier):
# x8 f64 flat narrow streams, band x32 f32 row-block with two tiles per lane, the shipped band substitution kernels; launch list of the bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_kernel -s 4 -c 1 -o gpurun_out/r2_prof_rows_l3d_n8_flat -f python tools/sweep.py --workload laplace3d_256_n8_f64 --steps 2 --warmup 2 --points ";" > gpurun_out/r2c23_ncu1.log 2>&1; echo "ncu n8 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rowblock_kernel -s 4 -c 1 -o gpurun_out/r2_prof_rowblock_band_n32_tiles2 -f python tools/sweep.py --workload band_1m_hb32_n32_f32 --steps 2 --warmup 2 --points ";" > gpurun_out/r2c23_ncu2.log 2>&1; echo "ncu rowblock rc=$?"
cat > gpurun_out/solve_small.py <<'PY'
import sys, os; sys.path.insert(0, os.getcwd()); import json, torch, bench
from basic_sparse_matrix_b200 import gpu
torch.cuda.set_device(0); gpu.init(0)
st = torch.cuda.Stream(); gpu.set_stream(st.cuda_stream); torch.cuda.set_stream(st)
print(json.dumps(bench.run_solve(torch, gpu, n_rows=1 << 18)))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trisolve_band -s 2 -c 2 -o gpurun_out/r2_prof_trisolve_band_final -f python gpurun_out/solve_small.py > gpurun_out/r2c23_ncu3.log 2>&1; echo "ncu solve rc=$?"
for f in r2_prof_rows_l3d_n8_flat r2_prof_rowblock_band_n32_tiles2 r2_prof_trisolve_band_final; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/${f}_raw.csv > gpurun_out/${f}_summary.txt 2>&1; rm -f gpurun_out/$f.ncu-rep; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2c23_ncu_launch.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/r2_launches_bench_final.csv
