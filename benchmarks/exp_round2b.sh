set -u
bash benchmarks/round_check.sh
timeout 300 python benchmarks/micro.py c4 --iters 6 > gpurun_out/r2c_micro_c4.jsonl 2> gpurun_out/r2c_micro_c4.err; cat gpurun_out/r2c_micro_c4.jsonl; tail -3 gpurun_out/r2c_micro_c4.err
