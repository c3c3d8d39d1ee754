#!/bin/bash
# Ablation build: tools/build_variant.sh <name> <nvcc -D flags...>  ->  powspec_b200/_build/lib_<name>.so
# (assign_tiles.cu recompiled with the flags, the other objects of the regular build reused);
# run it with POWSPEC_B200_LIBRARY=powspec_b200/_build/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
B=powspec_b200/_build
nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I include "$@" \
  -c powspec_b200/csrc/assign_tiles.cu -o $B/assign_tiles_$name.o
objs=$(ls $B/*.o | grep -v "assign_tiles")
nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $B/lib_$name.so $objs $B/assign_tiles_$name.o \
  -L /usr/local/cuda/lib64 -lcufft -Xlinker -rpath -Xlinker /usr/local/cuda/lib64 -Xlinker -Bsymbolic
echo $B/lib_$name.so
