#!/usr/bin/env bash
# One GPU-box pass over everything the round is judged on: the GPU parity tests, smoke(), the headline bench line,
# the transform micro-benchmark and one ncu capture of the transform kernel.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 450 python -m pytest tests -q -m gpu > gpurun_out/round_tests.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/round_tests.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/round_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/round_smoke.log
timeout 200 python bench.py > gpurun_out/round_bench.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/round_bench.log > gpurun_out/round_bench.json; cut -c1-400 gpurun_out/round_bench.json
timeout 90 python benchmarks/micro.py transform > gpurun_out/round_micro_transform.jsonl 2>gpurun_out/round_micro_transform.err; echo "micro rc=$?"; cut -c1-220 gpurun_out/round_micro_transform.jsonl
timeout 120 ncu --set full --clock-control none --import-source on -k regex:transform_kernel -c 7 -o gpurun_out/prof_transform -f \
  python benchmarks/micro.py transform --n 4000000 --iters 1 > gpurun_out/round_ncu_transform.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -n 8
