#!/bin/bash
# seventh GPU call: parity after the gather-engine refactor (vectorised A reads, merge kernel on the engine), sweeps, bench
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_g.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_g.log
tail -3 gpurun_out/gpu_tests_g.log
P128=";reg_flavour=2;reg_flavour=1;reg_flavour=4;rows_per_slice=8;rows_per_slice=32;stages=2;rows_per_warp=128;rows_per_warp=512;warps_per_cta=4,reg_flavour=3,ctas_per_sm=6"
timeout 900 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 8 --points "$P128" --out gpurun_out/sweepg_l3d_n128.jsonl > gpurun_out/sweepg_l3d_n128.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 8 --points "$P128" --out gpurun_out/sweepg_l3d_n64.jsonl > gpurun_out/sweepg_l3d_n64.log 2>&1
PM=";merge_items=128;merge_items=512;merge_items=384;warps_per_cta=4;warps_per_cta=4,merge_items=512;col_tile=32"
timeout 600 python tools/sweep.py --workload rmat20_n64_f64 --algo merge --steps 8 --points "$PM" --out gpurun_out/sweepg_rmat_f64.jsonl > gpurun_out/sweepg_rmat_f64.log 2>&1
timeout 600 python tools/sweep.py --workload rmat20_n64_f32 --algo merge --steps 8 --points "$PM" --out gpurun_out/sweepg_rmat_f32.jsonl > gpurun_out/sweepg_rmat_f32.log 2>&1
timeout 600 python tools/sweep.py --workload rmat20_n64_f64 --algo vector --steps 3 --points ";reg_flavour=1" --out gpurun_out/sweepg_rmat_f64_vector.jsonl > gpurun_out/sweepg_rmat_f64_vector.log 2>&1
timeout 1500 python bench.py > gpurun_out/bench_full_g.json 2> gpurun_out/bench_full_g.err; echo "bench exit $?" >> gpurun_out/bench_full_g.err
CMD="python tools/sweep.py --workload rmat20_n64_f64 --steps 2 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_grmat.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_g_rmat.csv $CMD > gpurun_out/ncu_launches_g.log 2>&1
echo done
