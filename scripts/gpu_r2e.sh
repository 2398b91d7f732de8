#!/bin/bash
# round-2 GPU session E (1 GPU): tests with the one-kernel halo exchange / uniform transport, quickbench, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r2e_pytest.log 2>&1; tail -4 gpurun_out/r2e_pytest.log
for lib in nextsimdg_b200/libnsdg_cuda.so build/variants/tr_minb5.so; do
  for rheo in mevp bbm; do
    NSDG_CUDA_LIB=$lib QB_RHEO=$rheo python scripts/quickbench.py
    NSDG_CUDA_LIB=$lib QB_RHEO=$rheo QB_DISTORT=1 python scripts/quickbench.py
  done
done 2>&1 | tee gpurun_out/r2e_quickbench.txt
QB_RHEO=mevp timeout 600 ncu --set full --clock-control none -k regex:transport_stage -s 2 -c 1 -o /tmp/r2e_transport -f python scripts/quickbench.py > gpurun_out/r2e_ncu_transport.log 2>&1
ncu -i /tmp/r2e_transport.ncu-rep --page raw --csv > gpurun_out/r2e_transport.csv 2>/dev/null
