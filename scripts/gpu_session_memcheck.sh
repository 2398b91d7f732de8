#!/bin/bash
# single-GPU sanitizer session: compute-sanitizer memcheck over the kernels new in this round, then tests + quickbench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -q -x \
  "tests/test_gpu_kat.py::test_rk_orders_match_the_oracle_step_by_step" "tests/test_gpu_kat.py::test_periodic_seam_in_the_module_path" \
  "tests/test_gpu_parity.py::test_single_subcycle" "tests/test_gpu_parity.py::test_keep_dg_moments_extension" \
  "tests/test_reference_vectors.py" "tests/test_gpu_parity.py::test_ragged_sizes" > gpurun_out/r2_san_memcheck.log 2>&1
echo "memcheck exit code $?" | tee -a gpurun_out/r2_san_memcheck.log; grep -c "Invalid\|out of bounds" gpurun_out/r2_san_memcheck.log; tail -6 gpurun_out/r2_san_memcheck.log
(python -m pytest tests -m gpu -q -x) > gpurun_out/r2_san_pytest.log 2>&1; tail -3 gpurun_out/r2_san_pytest.log
for rheo in mevp bbm; do QB_RHEO=$rheo python scripts/quickbench.py; done
