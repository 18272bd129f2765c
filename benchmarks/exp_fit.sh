set -u
mkdir -p gpurun_out
timeout 200 python benchmarks/fit_probe.py > gpurun_out/r2_micro_fit.jsonl 2> gpurun_out/r2_micro_fit.err; cat gpurun_out/r2_micro_fit.jsonl; tail -3 gpurun_out/r2_micro_fit.err
