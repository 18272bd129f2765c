set -u
bash benchmarks/round_check.sh
timeout 300 python benchmarks/micro.py c4 --iters 6 > gpurun_out/r2c_micro_c4.jsonl 2> gpurun_out/r2c_micro_c4.err; cat gpurun_out/r2c_micro_c4.jsonl; tail -3 gpurun_out/r2c_micro_c4.err
timeout 200 python benchmarks/micro.py sweep --sorted-idx --iters 3 --only rosenbrock 2>&1 | tail -1 | cut -c1-250
