#!/bin/bash
# final single-GPU record of the round: full test suite, the full default bench line, the reference arm, the ncu launch list of a
# short bench run, memcheck over the TMA-staged kernels, smoke
cd "$(dirname "$0")/.."
tag=${1:-r2n}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
(time timeout 900 python bench.py) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; head -c 600 gpurun_out/${tag}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; head -c 300 gpurun_out/${tag}_bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-parity-check --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_launches_bench.log 2>&1
TESTS="tests/test_gpu_parity.py::test_single_subcycle tests/test_gpu_parity.py::test_ragged_sizes tests/test_gpu_parity.py::test_parametric_factored_path_equals_streamed_operator_path"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -q -x $TESTS > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck exit code $?" | tee -a gpurun_out/${tag}_memcheck.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/${tag}_memcheck.log | tail -3
(time python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -6
du -sh gpurun_out
