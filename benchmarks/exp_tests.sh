set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2g_tests.log
timeout 200 python benchmarks/micro.py flow > gpurun_out/r2g_micro_flow.jsonl 2>&1; cut -c1-330 gpurun_out/r2g_micro_flow.jsonl
timeout 300 python benchmarks/micro.py c4 --iters 6 > gpurun_out/r2c_micro_c4.jsonl 2> gpurun_out/r2c_micro_c4.err; cat gpurun_out/r2c_micro_c4.jsonl; tail -3 gpurun_out/r2c_micro_c4.err
timeout 200 python benchmarks/micro.py importance > gpurun_out/r2g_micro_importance.jsonl 2>&1; cut -c1-300 gpurun_out/r2g_micro_importance.jsonl
