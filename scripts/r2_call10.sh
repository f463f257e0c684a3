#!/bin/bash
# Round 2, GPU call 10: plane-blocked super-batches (the warps of a CTA = planes x lines instead of lines of one plane) and a CTA
# barrier per slice, on the stencil workloads. Model: L2->L1 fills per output row = 1 + 2/Ly + 2/Lz (3.25 today, Ly=8, Lz=1).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
P128=";planes=2;planes=4;planes=8;sync=1;planes=2,sync=1;reg_flavour=7;reg_flavour=7,planes=3;reg_flavour=7,planes=6;reg_flavour=7,planes=3,sync=1;reg_flavour=7,sync=1;reg_flavour=1;reg_flavour=1,planes=4;reg_flavour=1,planes=2;reg_flavour=1,planes=4,sync=1;"
timeout 400 python tools/sweep.py --workload laplace3d_256_n128_f64 --steps 10 --points "$P128" --out gpurun_out/r2_sweep_planes_l3d_n128.jsonl > gpurun_out/r2c10_n128.log 2>&1; echo "n128 rc=$?"
python tools/show_sweep.py gpurun_out/r2_sweep_planes_l3d_n128.jsonl 2>/dev/null | cut -c1-150
P64=";planes=3;planes=6;planes=3,sync=1;sync=1;planes=6,sync=1;reg_flavour=5;reg_flavour=5,planes=2;reg_flavour=5,planes=4;"
timeout 400 python tools/sweep.py --workload laplace3d_256_n64_f64 --steps 10 --points "$P64" --out gpurun_out/r2_sweep_planes_l3d_n64.jsonl > gpurun_out/r2c10_n64.log 2>&1; echo "n64 rc=$?"
python tools/show_sweep.py gpurun_out/r2_sweep_planes_l3d_n64.jsonl 2>/dev/null | cut -c1-150
