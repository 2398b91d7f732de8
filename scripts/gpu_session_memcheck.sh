#!/bin/bash
# single-GPU sanitizer session: compute-sanitizer memcheck / initcheck / synccheck over the kernels new in round 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TESTS="tests/test_gpu_kat.py::test_rk_orders_match_the_oracle_step_by_step tests/test_gpu_kat.py::test_periodic_seam_in_the_module_path tests/test_gpu_parity.py::test_single_subcycle tests/test_gpu_parity.py::test_keep_dg_moments_extension tests/test_reference_vectors.py tests/test_gpu_parity.py::test_ragged_sizes tests/test_paragrid.py"
for tool in memcheck initcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest -q -x $TESTS > gpurun_out/r2_san_$tool.log 2>&1
  echo "$tool exit code $?" | tee -a gpurun_out/r2_san_$tool.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r2_san_$tool.log | tail -3
done
