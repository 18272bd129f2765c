#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_flow.py -x -q 2>&1 | tail -3
timeout 200 python benchmarks/micro.py flow > gpurun_out/r2e_micro_flow.jsonl 2>&1; cut -c1-330 gpurun_out/r2e_micro_flow.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flow_kernel -s 2 -c 1 -o gpurun_out/r2e_flow -f \
  python benchmarks/micro.py flow --n 4000000 --iters 1 > gpurun_out/r2e_ncu_flow.log 2>&1; echo "ncu rc=$?"
