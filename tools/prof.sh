#!/bin/bash
# on the GPU box: one ncu --set full capture of a mid-depth k_bounce launch of a short bench run, exported as CSV
# usage: tools/prof.sh <tag> [skip] [extra bench args]   -> gpurun_out/<tag>_{raw,src}.csv, <tag>_prof.ncu-rep
tag=$1; skip=${2:-9}; shift; shift
python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu-baseline "$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_bounce -s $skip -c 1 -o gpurun_out/${tag}_prof -f \
  python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_src.csv
