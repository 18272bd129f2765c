#!/usr/bin/env bash
# Round-2 sweep experiment: fixed part of the sweep (gather, initial target, final logs, store) vs moves, random vs sorted ancestors
set -u
mkdir -p gpurun_out
run() { # moves extra tag
  timeout 120 python benchmarks/micro.py sweep --iters 3 --only rosenbrock --moves $1 $2 2>&1 | grep '^{' | python -c "import sys,json; [print('$3 moves $1', round(json.loads(l)['ms'],3), 'ms acc', round(json.loads(l)['acceptance'],5)) for l in sys.stdin]"
}
for mv in 0 1 2 10; do run $mv "" random; run $mv "--sorted-idx" sorted; done
timeout 300 python -m pytest tests/test_gpu_sorted.py -x -q 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"search_sorted|spacings|tile_scan" -c 12 python -c "
import torch
from aspire_b200 import ops
n=100_000_000
lw = ops.philox_normal(3, 1, n, 1, 'cuda').reshape(-1) * 2.0
p = torch.exp(lw - lw.max()); p /= p.sum()
cdf = ops.cdf_exact(p)
for i in range(4):
    idx = ops.search_multinomial_sorted(cdf, 17, i, n)
torch.cuda.synchronize()
" 2>&1 | grep -E "asp::|gpu__time" | paste - - | awk '{print $1, $2, $NF}' | tail -9
