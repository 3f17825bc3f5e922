#!/bin/bash
# Build an experimental variant of libhopgpu.so with extra nvcc flags:  tools/build_variant.sh <name> <flags...>
# -> hevc-hop_b200/build/variants/libhopgpu_<name>.so ; select it with HOP_LIB=<path>
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/hevc-hop_b200/build/variants; mkdir -p $out/$name
for f in $root/hevc-hop_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=false "$@" -c $f -o $out/$name/$(basename $f .cu).o &
done
wait
nvcc -shared -arch=sm_100a -o $out/libhopgpu_$name.so $out/$name/*.o -cudart static
echo $out/libhopgpu_$name.so
