#!/bin/bash
# round-2 GPU session C: tests, transport variants, ncu --set full of the four fast strip kernels (DRAM traffic per launch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r2c_pytest.log 2>&1; tail -4 gpurun_out/r2c_pytest.log
for lib in "" build/variants/tr_minb3.so build/variants/tr_minb4.so; do
  for rheo in mevp bbm; do
    NSDG_CUDA_LIB=${lib:-nextsimdg_b200/libnsdg_cuda.so} QB_RHEO=$rheo python scripts/quickbench.py
    NSDG_CUDA_LIB=${lib:-nextsimdg_b200/libnsdg_cuda.so} QB_RHEO=$rheo QB_DISTORT=1 python scripts/quickbench.py
  done
done 2>&1 | tee gpurun_out/r2c_quickbench.txt
for cfg in "mevp:" "bbm:" "mevp:1" "bbm:1"; do
  rheo=${cfg%%:*}; dist=${cfg##*:}
  tag=${rheo}_$([ -n "$dist" ] && echo para || echo rect)
  QB_RHEO=$rheo QB_DISTORT=$dist timeout 600 ncu --set full --clock-control none -k regex:subcycle_strip -s 3 -c 1 \
     -o gpurun_out/r2c_strip_$tag -f python scripts/quickbench.py > gpurun_out/r2c_ncu_$tag.log 2>&1
  ncu -i gpurun_out/r2c_strip_$tag.ncu-rep --page raw --csv > gpurun_out/r2c_strip_$tag.csv 2>/dev/null
  tail -2 gpurun_out/r2c_ncu_$tag.log
done
ls -la gpurun_out | tail -12
