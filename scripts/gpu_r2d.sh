#!/bin/bash
# round-2 GPU session D: tests, transport variants, ncu --set full of the four fast strip kernels, CPU reference curve
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r2d_pytest.log 2>&1; tail -4 gpurun_out/r2d_pytest.log
for lib in nextsimdg_b200/libnsdg_cuda.so build/variants/tr_noshfl.so build/variants/tr_minb3.so build/variants/tr_minb5.so; do
  for rheo in mevp bbm; do
    NSDG_CUDA_LIB=$lib QB_RHEO=$rheo python scripts/quickbench.py
    NSDG_CUDA_LIB=$lib QB_RHEO=$rheo QB_DISTORT=1 python scripts/quickbench.py
  done
done 2>&1 | tee gpurun_out/r2d_quickbench.txt
for cfg in "mevp:" "bbm:" "mevp:1" "bbm:1"; do
  rheo=${cfg%%:*}; dist=${cfg##*:}
  tag=${rheo}_$([ -n "$dist" ] && echo para || echo rect)
  QB_RHEO=$rheo QB_DISTORT=$dist timeout 600 ncu --set full --clock-control none -k regex:subcycle_strip -s 3 -c 1 \
     -o /tmp/r2d_strip_$tag -f python scripts/quickbench.py > gpurun_out/r2d_ncu_$tag.log 2>&1
  ncu -i /tmp/r2d_strip_$tag.ncu-rep --page raw --csv > gpurun_out/r2d_strip_$tag.csv 2>/dev/null
  tail -1 gpurun_out/r2d_ncu_$tag.log
done
# the fused transport kernel under ncu (one launch, 2 fields)
QB_RHEO=mevp timeout 600 ncu --set full --clock-control none -k regex:transport_stage -s 2 -c 1 -o /tmp/r2d_transport -f python scripts/quickbench.py > gpurun_out/r2d_ncu_transport.log 2>&1
ncu -i /tmp/r2d_transport.ncu-rep --page raw --csv > gpurun_out/r2d_transport.csv 2>/dev/null
python scripts/cpu_reference_curve.py 2>&1 | tail -6
du -sh gpurun_out
