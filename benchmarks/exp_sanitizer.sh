set -u
mkdir -p gpurun_out
# memcheck of the kernels added in round 2b (flow sweep, fused tpCN, device flow fit, Student-t EM, NaN counters)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x \
  "tests/test_gpu_flow.py::test_flow_sweep_equals_the_semi_fused_moves" \
  "tests/test_gpu_flow.py::test_native_fit_gradients_and_updates_match_torch_autograd" \
  "tests/test_gpu_parity.py::test_fused_tpcn_sweep_equals_unfused_moves_philox" \
  "tests/test_gpu_parity.py::test_tpcn_moves_supplied_noise_match_oracle" \
  "tests/test_gpu_parity.py::test_sweep_counts_nan_log_q" \
  "tests/test_gpu_samplers.py::test_smc_with_tpcn_step" \
  > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r2_sanitizer_memcheck.log | head -20
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x \
  "tests/test_gpu_flow.py::test_flow_sweep_equals_the_semi_fused_moves" \
  "tests/test_gpu_parity.py::test_fused_tpcn_sweep_equals_unfused_moves_philox" \
  > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitizer_racecheck.log | head -10
