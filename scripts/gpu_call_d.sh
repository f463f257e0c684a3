#!/bin/bash
# fourth GPU call: parity, prefetch sweeps, first full bench line (+ reference arm), ncu launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host.py -m gpu -x -q > gpurun_out/gpu_tests_d.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_d.log
tail -3 gpurun_out/gpu_tests_d.log
P128=";prefetch_dist=-1;prefetch_dist=16;prefetch_dist=64;prefetch_dist=128;reg_flavour=2;reg_flavour=2,prefetch_dist=-1;reg_flavour=1;reg_flavour=1,prefetch_dist=-1;reg_flavour=4;rows_per_slice=8;rows_per_slice=32;stages=2;stages=4;col_tile=64;rows_per_warp=128;rows_per_warp=16"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 8 --points "$P128" --out gpurun_out/sweepd_l3d_n128.jsonl > gpurun_out/sweepd_l3d_n128.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweepd_l3d_n128.log
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 8 --points "$P128" --out gpurun_out/sweepd_l3d_n64.jsonl > gpurun_out/sweepd_l3d_n64.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweepd_l3d_n64.log
PS=";rows_per_slice=64;rows_per_slice=128;rows_per_slice=256;warps_per_cta=8;rows_per_warp=256;stages=2"
timeout 300 python tools/sweep.py --workload laplace2d_2048_n1_f64 --algo vector --steps 20 --points "$PS" --out gpurun_out/sweepd_l2d_n1.jsonl > gpurun_out/sweepd_l2d_n1.log 2>&1
PB=";rows_per_warp=32;rows_per_warp=64;rows_per_slice=8;rows_per_slice=32;warps_per_cta=8;stages=2"
timeout 300 python tools/sweep.py --workload band_1m_hb32_n32_f32 --algo vector --steps 20 --points "$PB" --out gpurun_out/sweepd_band_n32.jsonl > gpurun_out/sweepd_band_n32.log 2>&1
timeout 300 python tools/sweep.py --workload band_1m_hb32_n1_f32 --algo vector --steps 20 --points "$PB" --out gpurun_out/sweepd_band_n1.jsonl > gpurun_out/sweepd_band_n1.log 2>&1
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err
echo done
