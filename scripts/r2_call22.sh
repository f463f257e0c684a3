#!/bin/bash
# Round 2, GPU call 22: the last-entry "prefetch" as a REAL TMA load of the B row into a scratch nobody reads (an ordinary L2 fill, no prefetch semantics)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
P=";pf=9;pf=10;pf=11;pf=1;rows_per_slice=8,pf=9;rows_per_slice=8,pf=10;rows_per_slice=4,pf=10;"
for wl in laplace3d_256_n128_f64 laplace3d_256_n64_f64; do
timeout 400 python tools/sweep.py --workload $wl --steps 10 --points "$P" --out gpurun_out/r2_sweep_scratchload_$wl.jsonl > gpurun_out/r2c22_$wl.log 2>&1; echo "$wl rc=$?"
python tools/show_sweep.py gpurun_out/r2_sweep_scratchload_$wl.jsonl | cut -c1-150
done
