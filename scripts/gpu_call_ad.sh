#!/bin/bash
# does the grouped-lane default also hold on regular rows WITHOUT locality (uniformly random columns)?
set -u
mkdir -p gpurun_out
timeout 900 python tools/sweep.py --workload uniform22_n64_f64 --algo vector --steps 10 --points ";lanes_per_row=32;;lanes_per_row=32;lanes_per_row=16;lanes_per_row=8,reg_flavour=5" --out gpurun_out/sweepad_uniform_n64.jsonl > gpurun_out/sweepad_uniform_n64.log 2>&1
timeout 900 python tools/sweep.py --workload uniform22_n32_f64 --algo vector --steps 10 --points ";lanes_per_row=32;;lanes_per_row=32;lanes_per_row=4" --out gpurun_out/sweepad_uniform_n32.jsonl > gpurun_out/sweepad_uniform_n32.log 2>&1
timeout 900 python tools/sweep.py --workload uniform22_n64_f64 --algo merge --steps 10 --points ";lanes_per_row=32,merge_items=384" --out gpurun_out/sweepad_uniform_n64_merge.jsonl > gpurun_out/sweepad_uniform_n64_merge.log 2>&1
echo done
