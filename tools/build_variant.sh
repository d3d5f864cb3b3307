#!/bin/bash
# Experiment libraries: the shipped objects with ONE source recompiled under extra flags.
#   tools/build_variant.sh <name> <source.cu> <nvcc flags...>   ->  rust-compression_b200/build/variants/lib_<name>.so
# tools/gpu_ab.sh runs a variant given as "LIB=<name>" by copying that file over libbzb200.so on the GPU box's scratch copy.
set -e
cd "$(dirname "$0")/../rust-compression_b200"
name=$1; src=$2; shift 2
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
nvcc $FLAGS "$@" -c csrc/$src -o build/variants/${name}_${src%.cu}.o
objs=""
for f in k1_rle k2_bwt k3_mtf k4_huff k6_pack decoder pipeline slice_plan mgpu enc_stream dec_abi; do
  if [ "$f.cu" = "$src" ]; then objs="$objs build/variants/${name}_$f.o"; else objs="$objs build/$f.o"; fi
done
nvcc -shared -o build/variants/lib_$name.so $objs -gencode arch=compute_100a,code=sm_100a
echo build/variants/lib_$name.so
