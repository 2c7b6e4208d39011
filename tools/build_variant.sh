#!/bin/bash
# build an alternative libpt_b200 with extra -D flags:  tools/build_variant.sh name -DPT_MIN_BLOCKS=3 ...
# load it with PT_B200_LIB=build/libpt_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 \
  "$@" -I include -shared -o build/libpt_$name.so project3-pathtracer_b200/csrc/*.cu \
  $(ls project3-pathtracer_b200/csrc/*.cpp | grep -v pt_main.cpp) -lz
echo build/libpt_$name.so
