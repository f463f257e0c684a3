#!/bin/bash
# Round 2, GPU call 9: L2 prefetch of the B row of every row's LAST stored entry (the first touch in a row-order sweep of a
# stencil matrix), issued once per slice outside the entry loop. Sweep over mode bits and slice lengths, same box A/B.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
P128=";pf=1;pf=2;pf=3;pf=5;pf=6;rows_per_slice=8;rows_per_slice=8,pf=1;rows_per_slice=8,pf=2;rows_per_slice=8,pf=6;rows_per_slice=4,pf=2;rows_per_slice=32,pf=1;"
timeout 400 python tools/sweep.py --workload laplace3d_256_n128_f64 --steps 10 --points "$P128" --out gpurun_out/r2_sweep_pflast_l3d_n128.jsonl > gpurun_out/r2c9_n128.log 2>&1; echo "n128 rc=$?"
python tools/show_sweep.py gpurun_out/r2_sweep_pflast_l3d_n128.jsonl 2>/dev/null || cut -c1-160 gpurun_out/r2_sweep_pflast_l3d_n128.jsonl
timeout 400 python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 10 --points "$P128" --out gpurun_out/r2_sweep_pflast_l3d_n64.jsonl > gpurun_out/r2c9_n64.log 2>&1; echo "n64 rc=$?"
python tools/show_sweep.py gpurun_out/r2_sweep_pflast_l3d_n64.jsonl 2>/dev/null || cut -c1-160 gpurun_out/r2_sweep_pflast_l3d_n64.jsonl
timeout 400 python tools/sweep.py --workload laplace2d_4096_n64_f64 --steps 10 --points ";pf=1;pf=2;pf=3" --out gpurun_out/r2_sweep_pflast_l2d_n64.jsonl > gpurun_out/r2c9_l2d.log 2>&1; echo "l2d rc=$?"
python tools/show_sweep.py gpurun_out/r2_sweep_pflast_l2d_n64.jsonl 2>/dev/null
