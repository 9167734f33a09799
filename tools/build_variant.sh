#!/bin/bash
# A/B builds of the same ABI: tools/build_variant.sh <tag> <nvcc -D flags...>  -> devastator_b200/libdeva_b200_<tag>.so
# (run with QB_LIB=devastator_b200/libdeva_b200_<tag>.so python tools/tile_probe.py)
tag=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --compiler-options -fPIC -shared "$@" \
  devastator_b200/csrc/engine.cu -o devastator_b200/libdeva_b200_$tag.so -ldl -Xptxas -v 2>&1 | grep -A2 "k_tile_step.*PholdTest" | grep -E "registers|spill" | tr '\n' ' '
echo " <- $tag"
