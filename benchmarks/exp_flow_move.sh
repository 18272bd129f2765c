set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flow.py tests/test_gpu_parity.py -q -x > gpurun_out/r2c_flow_tests.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2c_flow_tests.log
timeout 200 python benchmarks/micro.py flowsweep --iters 3 > gpurun_out/r2c_micro_flowsweep.jsonl 2> gpurun_out/r2c_micro_flowsweep.err; cat gpurun_out/r2c_micro_flowsweep.jsonl; tail -3 gpurun_out/r2c_micro_flowsweep.err
timeout 200 python benchmarks/micro.py sweep --sorted-idx --iters 3 --only gmm 2>&1 | tail -1 | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:flow_move -s 1 -c 1 -o gpurun_out/r2c_flow_move -f \
  python benchmarks/micro.py flowsweep --n 2000000 --iters 1 > gpurun_out/r2c_ncu_flow_move.log 2>&1; echo "ncu rc=$?"
