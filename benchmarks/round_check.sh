#!/usr/bin/env bash
# One GPU-box pass: the GPU parity tests, smoke(), the headline bench line.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/round_tests.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/round_tests.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/round_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/round_smoke.log
timeout 300 python bench.py > gpurun_out/round_bench.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/round_bench.log > gpurun_out/round_bench.json; cut -c1-1200 gpurun_out/round_bench.json
