set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r2b_tests.log
