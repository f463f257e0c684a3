#!/bin/bash
# sanitizer call 1: compute-sanitizer memcheck over the small parity cases (one tool per call)
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file gpurun_out/memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_cpp_host.py -m gpu -x -q -k "reference_kats or mul_vector_kat or edge_shapes or tuning_variants or merge_items or zero_drop or slow_path or pipelined or single_giant or cpp_reference" > gpurun_out/memcheck_pytest.log 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck_pytest.log
tail -5 gpurun_out/memcheck_pytest.log
tail -5 gpurun_out/memcheck.log
echo done
