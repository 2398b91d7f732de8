#!/bin/bash
# build a variant of libnsdg_cuda.so with extra -D flags into build/variants/<name>.so (travels to the GPU box; git-ignored)
#   scripts/build_variant.sh name -DNSDG_FOO=1 ...
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -std=c++20 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --expt-relaxed-constexpr \
  -Xcompiler -fPIC -shared --threads 0 "$@" -o build/variants/$name.so nextsimdg_b200/csrc/nsdg_cuda.cu nextsimdg_b200/csrc/nsdg_kernels_*.cu
