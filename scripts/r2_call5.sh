#!/bin/bash
# Round 2, GPU call 5 (1 GPU): occupancy experiments. Each point in its own process (a faulting variant must not poison the rest).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { # workload tag points [extra args]
  timeout 200 python tools/sweep.py --steps 10 --algo vector --workload $1 --points "$3" ${4:-} --out gpurun_out/r2_occ_$1_$2.jsonl > gpurun_out/r2c5_$1_$2.log 2>&1
  echo "$1 $2: $(cut -c1-100 gpurun_out/r2_occ_$1_$2.jsonl | tr '\n' ' ')"
}
run laplace3d_256_n128_f64 default ";"
run laplace3d_256_n128_f64 f9 "reg_flavour=9"
run laplace3d_256_n128_f64 f9s2 "reg_flavour=9,stages=2"
run laplace3d_256_n128_f64 f9s2r8 "reg_flavour=9,stages=2,rows_per_slice=8"
run laplace3d_256_n128_f64 f10 "reg_flavour=10"
run laplace3d_256_n128_f64 f10r8 "reg_flavour=10,rows_per_slice=8"
run laplace3d_256_n128_f64 default2 ";"
run laplace3d_252_n128_f64 f9 ";reg_flavour=9"
run laplace3d_256_n128_f32 f9 ";reg_flavour=9"
for N in 4 8 16; do
  run laplace3d_256_n${N}_f64 default ";"
  run laplace3d_256_n${N}_f64 f9 "reg_flavour=9"
  run laplace3d_256_n${N}_f64 f9s2 "reg_flavour=9,stages=2"
  run laplace3d_256_n${N}_f64 f9r16 "reg_flavour=9,rows_per_slice=16"
done
run laplace3d_256_n32_f64 f9 ";reg_flavour=9"
run band_1m_hb32_n1_f32 default ";"
