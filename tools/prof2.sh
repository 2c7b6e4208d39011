#!/bin/bash
# on the GPU box: ncu --set full of ONE launch of kernel regex $2 (skipping $3 matches) in a short bench run
# usage: tools/prof2.sh <tag> <kernel-regex> [skip] [extra bench args]
tag=$1; rx=$2; skip=${3:-2}; shift; shift; shift
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/${tag}_prof -f \
  python bench.py --steps 1 --warmup 0 --spp 16 --wf-spp 16 --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_src.csv
