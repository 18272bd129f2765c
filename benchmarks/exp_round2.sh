#!/usr/bin/env bash
# Round-2 measurement pass on one GPU: C5 micro-benchmark with numpy beside it, ncu --set full of the sweep and of the
# sorted search, launch list of the default bench.
set -u
mkdir -p gpurun_out
timeout 600 python benchmarks/micro.py resample --iters 5 > gpurun_out/r2_micro_resample.jsonl 2> gpurun_out/r2_micro_resample.err; echo "resample rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_micro_resample.jsonl'):
    d=json.loads(l); print(d['bench'], d['particles'], round(d['ms_total'],3), 'ms', 'numpy', d.get('numpy_host_ms'), 'x', d.get('vs_numpy'))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcn_sweep_tpp -s 3 -c 1 -o gpurun_out/r2_sweep -f \
  python benchmarks/micro.py sweep --only rosenbrock --n 8000000 --iters 1 --sorted-idx > gpurun_out/r2_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"search_sorted|spacings_kernel|beta_node|beta_newton" -c 8 -o gpurun_out/r2_step_kernels -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --particles 20000000 > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu step rc=$?"
bash benchmarks/launch_list.sh > gpurun_out/r2_launches.md 2>&1; tail -30 gpurun_out/r2_launches.md
ls -la gpurun_out | tail -8
