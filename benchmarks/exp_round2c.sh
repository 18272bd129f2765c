set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_transforms.py tests/test_gpu_integration.py tests/test_gpu_parity.py -q -x > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2d_tests.log
timeout 120 python benchmarks/pcie_probe.py > gpurun_out/r2_pcie_probe.jsonl 2>&1; cat gpurun_out/r2_pcie_probe.jsonl
