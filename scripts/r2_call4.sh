#!/bin/bash
# Round 2, GPU call 4 (1 GPU): experiment — more resident warps with shallower gather windows on the headline kernel
# (reg_flavour 9: 4 CTAs x 8 warps, window 3; 10: 5 CTAs, window 2), and timing + DRAM bytes of the conversion /
# result-construction kernels inside one end-to-end call (ncu, duration and dram bytes only).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S="timeout 300 python tools/sweep.py --steps 10"
P=";reg_flavour=9;reg_flavour=9,stages=2;reg_flavour=9,stages=2,rows_per_slice=8;reg_flavour=10;reg_flavour=10,rows_per_slice=8;reg_flavour=9,rows_per_warp=128;"
$S --workload laplace3d_256_n128_f64 --algo vector --points "$P" --out gpurun_out/r2_sweep_occupancy_l3d_n128.jsonl > gpurun_out/r2c4_s1.log 2>&1
$S --workload laplace3d_252_n128_f64 --algo vector --points ";reg_flavour=9;reg_flavour=9,stages=2;reg_flavour=10" --out gpurun_out/r2_sweep_occupancy_l3d252_n128.jsonl > gpurun_out/r2c4_s2.log 2>&1
$S --workload laplace3d_256_n128_f64 --algo vector --slice 3/8 --points ";reg_flavour=9;reg_flavour=9,stages=2;reg_flavour=10" --out gpurun_out/r2_sweep_occupancy_l3d_n128_s3_8.jsonl > gpurun_out/r2c4_s3.log 2>&1
for f in gpurun_out/r2_sweep_occupancy_*.jsonl; do echo $f; cut -c1-110 $f; done
timeout 600 python bench.py --workload laplace3d_256_n64_f64 --steps 3 --warmup 3 --no-extras --no-cpu --no-target --e2e-steps 1 > gpurun_out/r2c4_e2e_n64.json 2> gpurun_out/r2c4_e2e_n64.err
echo "plain e2e run rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"transpose_kernel|narrow_u64|widen_u32|count_nonzero|scatter_nonzero|scan_tile|scan_add|csr_stats|row_index_piece|block_col_range" -c 400 --csv \
  --log-file gpurun_out/r2_convert_kernels_e2e_n64.csv python bench.py --workload laplace3d_256_n64_f64 --steps 3 --warmup 3 --no-extras --no-cpu --no-target --e2e-steps 1 > gpurun_out/r2c4_ncu.log 2>&1
echo "ncu convert rc=$?"
wc -l gpurun_out/r2_convert_kernels_e2e_n64.csv
