#!/bin/bash
# usage: scratch/build_variant.sh <name> <nvcc -D flags...>  -> scratch/variants/librf_<name>.so (A/B builds; select with RF_B200_LIB)
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch/variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o scratch/variants/librf_$name.so riffusion-hobby_b200/csrc/*.cu riffusion-hobby_b200/csrc/*.cpp
echo built scratch/variants/librf_$name.so
