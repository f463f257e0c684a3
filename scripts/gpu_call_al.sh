#!/bin/bash
# upper bound of near-diagonal B-row sharing on the headline matrix (see tools/probe_near_diag.py)
set -u
mkdir -p gpurun_out
timeout 900 python tools/probe_near_diag.py > gpurun_out/probe_near_diag.jsonl 2> gpurun_out/probe_near_diag.err
cat gpurun_out/probe_near_diag.jsonl | cut -c1-150
tail -2 gpurun_out/probe_near_diag.err
echo done
timeout 600 python tools/sweep.py --workload laplace2d_4096_n64_f64 --algo vector --steps 10 --points ";rows_per_slice=24;rows_per_slice=32;rows_per_slice=16;" --out gpurun_out/sweepal_l2d_n64.jsonl > gpurun_out/sweepal_l2d_n64.log 2>&1
tail -4 gpurun_out/sweepal_l2d_n64.log | cut -c1-330
