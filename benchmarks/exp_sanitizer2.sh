set -u
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest -q -x \
  "tests/test_gpu_flow.py::test_flow_sweep_equals_the_semi_fused_moves" \
  "tests/test_gpu_flow.py::test_flow_sweep_with_the_independence_proposal" \
  "tests/test_gpu_samplers.py::test_smc_with_tpcn_step" \
  "tests/test_gpu_parity.py::test_weighted_moments_strided_rows_in_place" \
  > gpurun_out/r2_sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Uninitialized" gpurun_out/r2_sanitizer_initcheck.log | head -12
timeout 120 python examples/flow_proposal.py 2>&1 | tail -4
