set -u
mkdir -p gpurun_out
bash benchmarks/launch_list.sh > gpurun_out/r2_launches_final.md 2>&1; tail -45 gpurun_out/r2_launches_final.md
timeout 300 ncu --set full --clock-control none -k regex:pcn_sweep_tpp -s 1 -c 1 -o gpurun_out/r2_sweep_final -f \
  python benchmarks/micro.py sweep --only rosenbrock --n 8000000 --sorted-idx --iters 1 > gpurun_out/r2_ncu_sweep_final.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:pcn_sweep_tpp -s 1 -c 1 -o gpurun_out/r2_tpcn_final -f \
  python benchmarks/micro.py sweep --only rosenbrock --n 8000000 --sorted-idx --iters 1 --nu 5 > gpurun_out/r2_ncu_tpcn_final.log 2>&1; echo "ncu rc=$?"
