#!/usr/bin/env bash
# usage: exp_multi2.sh G   (run on a G-GPU box): multi-GPU tests (G = 2 only) and the headline bench at G ranks with the host-buffer leg
set -u
g=$1
mkdir -p gpurun_out
if [ "$g" = "2" ]; then
  timeout 400 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
fi
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g --steps 10 --warmup 3 > gpurun_out/bench_${g}gpu.log 2>&1
grep "^{" gpurun_out/bench_${g}gpu.log > gpurun_out/r2f_bench_${g}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_${g}gpu.json')); print($g, round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phase_ms'].items()}, d['multi_gpu_parity']['ok'], d['multi_gpu_parity']['max_abs_dx'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1))
o=d.get('other_configs') or {}
for l in o.get('C5_resample_sharded', []): print(l['bench'], l['particles'], round(l['ms_total'],3))
if 'error' in o: print(o)
" || tail -20 gpurun_out/bench_${g}gpu.log
