set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2g_tests.log
timeout 120 python examples/flow_proposal.py 2>&1 | tail -5
