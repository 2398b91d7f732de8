#!/bin/bash
# build a variant of libnsdg_cuda.so that differs from the main build in ONE translation unit (objects of the others are cached in
# build/obj): scripts/build_variant_fast.sh name unit -DNSDG_FOO=1 ...   (unit: umevp | ubbm | pmevp | pbbm | umevp1 | cuda)
cd "$(dirname "$0")/.."
name=$1; unit=$2; shift 2
mkdir -p build/variants build/obj
FLAGS="-std=c++20 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC"
for u in cuda kernels_umevp kernels_ubbm kernels_pmevp kernels_pbbm kernels_umevp1; do
  src=nextsimdg_b200/csrc/nsdg_$u.cu; obj=build/obj/$u.o
  if [ ! -f $obj ] || [ -n "$(find nextsimdg_b200/csrc include -newer $obj -name '*.cu*' -o -newer $obj -name '*.h' | head -1)" ]; then
    /usr/local/cuda/bin/nvcc $FLAGS -c -o $obj $src &
  fi
done
wait
vobj=build/obj/variant_$name.o
u=$unit; [ $u != cuda ] && u=kernels_$u
/usr/local/cuda/bin/nvcc $FLAGS "$@" -c -o $vobj nextsimdg_b200/csrc/nsdg_$u.cu || exit 1
objs=""
for x in cuda kernels_umevp kernels_ubbm kernels_pmevp kernels_pbbm kernels_umevp1; do
  if [ $x = $u ]; then objs="$objs $vobj"; else objs="$objs build/obj/$x.o"; fi
done
/usr/local/cuda/bin/nvcc -shared -o build/variants/$name.so $objs && echo built build/variants/$name.so
