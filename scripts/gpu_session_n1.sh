#!/bin/bash
# single-GPU session: full test suite, the full default bench line, ncu launch list of a short bench run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_n1_pytest.log 2>&1; tail -6 gpurun_out/r2_n1_pytest.log
(time python bench.py) > gpurun_out/r2_n1_bench.json 2> gpurun_out/r2_n1_bench.err; tail -3 gpurun_out/r2_n1_bench.err; head -c 400 gpurun_out/r2_n1_bench.json; echo
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2_n1_bench_reference.json 2> gpurun_out/r2_n1_bench_reference.err; head -c 300 gpurun_out/r2_n1_bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_n1_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-parity-check --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_n1_launches_bench.log 2>&1
(time python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -8
du -sh gpurun_out
