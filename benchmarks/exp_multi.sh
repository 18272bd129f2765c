#!/usr/bin/env bash
# Multi-GPU pass on an 8-GPU box: headline bench at 8 / 4 / 2 ranks, C5 resampling micro-benchmark sharded at 2 / 4 / 8.
set -u
mkdir -p gpurun_out
for g in 8 4 2; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_${g}gpu.log 2>&1
  grep "^{" gpurun_out/bench_${g}gpu.log > gpurun_out/bench_${g}gpu.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_${g}gpu.json')); print($g, round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phase_ms'].items()}, d['multi_gpu_parity']['ok'], d['multi_gpu_parity']['max_abs_dx'])"
done
for g in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29534 benchmarks/micro.py resample --gpus $g --iters 5 2> gpurun_out/micro_resample_${g}gpu.err | grep '^{' > gpurun_out/r2_micro_resample_${g}gpu.jsonl
  python - <<PY
import json
for l in open('gpurun_out/r2_micro_resample_${g}gpu.jsonl'):
    d=json.loads(l); print($g, d['bench'], d['particles'], round(d['ms_total'],3), 'ms cdf', round(d['ms_cdf'],3))
PY
done
