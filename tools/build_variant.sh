#!/bin/bash
# developer A/B: tools/build_variant.sh <name> <file.cu> <extra nvcc flags...>  ->  experiments/lib/libaspire_b200_<name>.so
# (the named source recompiled with the flags, every other object taken from the in-tree build)
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/.."
mkdir -p experiments/lib
obj=experiments/lib/${name}_$(basename ${src%.cu}).o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c aspire_b200/csrc/$src -o $obj
others=$(ls aspire_b200/csrc/build/*.o | grep -v "/$(echo ${src%.cu} | tr / _).o")
/usr/local/cuda/bin/nvcc -shared -o experiments/lib/libaspire_b200_${name}.so $obj $others -lcudart
echo experiments/lib/libaspire_b200_${name}.so
