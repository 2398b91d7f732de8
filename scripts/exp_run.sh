#!/bin/bash
# run on the GPU box: parity check + strip-kernel times of the library variants build/variants/<name>.so
#   QB_RHEO=bbm [QB_DISTORT=1] scripts/exp_run.sh name ...     (main = the product library)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  lib=build/variants/$v.so; [ "$v" = main ] && lib=nextsimdg_b200/libnsdg_cuda.so
  NSDG_CUDA_LIB=$PWD/$lib timeout 300 python scripts/check_variant.py 2>&1 | tail -1
  NSDG_CUDA_LIB=$PWD/$lib QB_RHEO=${QB_RHEO:-bbm} timeout 300 python scripts/quickbench.py 2>&1 | tail -1
done | tee -a gpurun_out/exp_run.log
