set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flow.py -q -x > gpurun_out/r2c_flow_tests.log 2>&1; echo "pytest rc=$?"; tail -n 30 gpurun_out/r2c_flow_tests.log
timeout 300 python benchmarks/micro.py c4 --iters 6 > gpurun_out/r2c_micro_c4.jsonl 2> gpurun_out/r2c_micro_c4.err; cat gpurun_out/r2c_micro_c4.jsonl; tail -5 gpurun_out/r2c_micro_c4.err
