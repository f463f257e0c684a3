#!/bin/bash
# Round 2, GPU call 1 (1 GPU): the whole GPU test-suite, then same-box A/B sweeps of this round's new knobs
# (L2 prefetch distance of the vector kernel, merge-path residency / carve-out, opt-in fused row-block arithmetic),
# then a short bench run. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log )
tail -5 gpurun_out/r2c1_pytest.log
S="timeout 400 python tools/sweep.py --steps 10"
PF=";b_prefetch=1;b_prefetch=2;b_prefetch=4;b_prefetch=8;b_prefetch=16;b_prefetch=32;b_prefetch=64;b_prefetch=128"
$S --workload laplace3d_256_n128_f64 --points "$PF" --out gpurun_out/r2_sweep_prefetch_l3d_n128.jsonl > gpurun_out/r2c1_s1.log 2>&1
$S --workload laplace3d_256_n64_f64 --points "$PF" --out gpurun_out/r2_sweep_prefetch_l3d_n64.jsonl > gpurun_out/r2c1_s2.log 2>&1
$S --workload laplace3d_256_n32_f64 --points ";b_prefetch=2;b_prefetch=4;b_prefetch=8;b_prefetch=16" --out gpurun_out/r2_sweep_prefetch_l3d_n32.jsonl > gpurun_out/r2c1_s3.log 2>&1
$S --workload laplace3d_256_n16_f64 --points ";b_prefetch=2;b_prefetch=4;b_prefetch=8;b_prefetch=16" --out gpurun_out/r2_sweep_prefetch_l3d_n16.jsonl > gpurun_out/r2c1_s4.log 2>&1
$S --workload laplace2d_4096_n64_f64 --points ";b_prefetch=4;b_prefetch=16;b_prefetch=64" --out gpurun_out/r2_sweep_prefetch_l2d_n64.jsonl > gpurun_out/r2c1_s5.log 2>&1
$S --workload laplace3d_252_n128_f64 --points ";b_prefetch=4;b_prefetch=16" --out gpurun_out/r2_sweep_prefetch_l3d252_n128.jsonl > gpurun_out/r2c1_s6.log 2>&1
$S --workload laplace3d_256_n128_f64 --slice 3/8 --points ";b_prefetch=4;b_prefetch=16" --out gpurun_out/r2_sweep_prefetch_l3d_n128_s3_8.jsonl > gpurun_out/r2c1_s7.log 2>&1
MG=";ctas_per_sm=1;ctas_per_sm=2;merge_items=96;merge_items=96,ctas_per_sm=2;merge_items=128;merge_items=128,ctas_per_sm=2;merge_items=256,ctas_per_sm=2;merge_items=384,ctas_per_sm=1;col_tile=32;col_tile=32,ctas_per_sm=2;lanes_per_row=8,col_tile=32"
$S --workload rmat20_n64_f64 --points "$MG" --out gpurun_out/r2_sweep_merge_rmat_f64.jsonl > gpurun_out/r2c1_s8.log 2>&1
$S --workload rmat20_n64_f32 --points ";ctas_per_sm=1;ctas_per_sm=2;merge_items=96;merge_items=96,ctas_per_sm=2;merge_items=256,ctas_per_sm=2;col_tile=32" --out gpurun_out/r2_sweep_merge_rmat_f32.jsonl > gpurun_out/r2c1_s9.log 2>&1
FU=";flags=7;flags=7,rows_per_slice=8;rows_per_slice=8"
$S --workload band_1m_hb32_n32_f32 --points "$FU" --out gpurun_out/r2_sweep_fused_band_n32.jsonl > gpurun_out/r2c1_s10.log 2>&1
$S --workload band_1m_hb32_n64_f64 --points "$FU" --out gpurun_out/r2_sweep_fused_band_n64_f64.jsonl > gpurun_out/r2c1_s11.log 2>&1
$S --workload band_1m_hb32_n128_f32 --points "$FU" --out gpurun_out/r2_sweep_fused_band_n128_f32.jsonl > gpurun_out/r2c1_s12.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2c1_bench.err
head -c 3000 gpurun_out/r2c1_bench.json
cat gpurun_out/r2_sweep_prefetch_l3d_n128.jsonl | cut -c1-120
