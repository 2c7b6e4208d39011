#!/bin/bash
# registers / spills of every k_bounce* variant:  tools/spills.sh [extra nvcc -D flags]
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 -Xptxas -v "$@" \
  -I include -c -o /tmp/pt_api_spills.o project3-pathtracer_b200/csrc/pt_api.cu 2>&1 | \
  awk '/Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)} /bytes stack frame/ {st=$0} /Used [0-9]+ registers/ {if (name ~ /k_bounce|k_bvh|k_shadow|k_raygen/) {gsub(/ptxas info +: /,"",st); gsub(/ptxas info +: /,"",$0); printf "%-62s %s | %s\n", name, st, $0}}' | sed 's/_ZN3ptd//; s/EvNS_12BounceParamsE//' | sort
