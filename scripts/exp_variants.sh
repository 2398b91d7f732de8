#!/bin/bash
# experiment: strip-kernel times of library variants under variants/ (run on the GPU box)
#   scripts/exp_variants.sh [-c] A B main ...      (-c: also run the parity check of each variant)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
chk=0; [ "$1" = "-c" ] && { chk=1; shift; }
for v in "$@"; do
  lib=variants/lib$v.so; [ "$v" = main ] && lib=nextsimdg_b200/libnsdg_cuda.so
  [ $chk = 1 ] && NSDG_CUDA_LIB=$PWD/$lib timeout 300 python scripts/check_variant.py 2>&1 | tail -1
  NSDG_CUDA_LIB=$PWD/$lib QB_RHEO=${QB_RHEO:-bbm} timeout 300 python scripts/quickbench.py 2>&1 | tail -1
done | tee -a gpurun_out/exp_variants.log
