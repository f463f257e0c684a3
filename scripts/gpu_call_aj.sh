#!/bin/bash
# tail split against wave quantisation: parity, then same-box A/B on 1/8, 1/4, 1/2 row blocks of the headline matrix
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail_split or tuning or random_shape" > gpurun_out/gpu_tests_aj.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/gpu_tests_aj.log
tail -2 gpurun_out/gpu_tests_aj.log
for sl in 3/8 1/4 1/2; do
  tag=$(echo $sl | tr / _)
  timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 20 --slice $sl --points ";flags=7;;flags=7" --out gpurun_out/sweepaj_l3d_n128_s$tag.jsonl > gpurun_out/sweepaj_l3d_n128_s$tag.log 2>&1
done
timeout 600 python tools/sweep.py --workload laplace3d_256_n64_f64 --algo vector --steps 20 --slice 3/8 --points ";flags=7;;flags=7" --out gpurun_out/sweepaj_l3d_n64_s3_8.jsonl > gpurun_out/sweepaj_l3d_n64_s3_8.log 2>&1
timeout 600 python tools/sweep.py --workload laplace3d_256_n128_f64 --algo vector --steps 10 --points ";flags=7;" --out gpurun_out/sweepaj_l3d_n128_full.jsonl > gpurun_out/sweepaj_l3d_n128_full.log 2>&1
echo done
