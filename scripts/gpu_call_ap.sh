#!/bin/bash
# 27-point box stencil: line detection nx-1, nx, nx+1 -> nx
set -u
mkdir -p gpurun_out
timeout 900 python tools/probe_box_stencil.py > gpurun_out/probe_box_stencil.jsonl 2> gpurun_out/probe_box_stencil.err
cut -c1-400 gpurun_out/probe_box_stencil.jsonl; tail -2 gpurun_out/probe_box_stencil.err
echo done
