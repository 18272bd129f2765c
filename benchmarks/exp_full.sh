set -u
mkdir -p gpurun_out
timeout 300 python benchmarks/micro.py full --n 100000000 > gpurun_out/r2_micro_full.jsonl 2> gpurun_out/r2_micro_full.err; tail -3 gpurun_out/r2_micro_full.err
timeout 300 python benchmarks/micro.py full --n 100000000 --only pcn >> gpurun_out/r2_micro_full.jsonl 2>> gpurun_out/r2_micro_full.err
cat gpurun_out/r2_micro_full.jsonl
