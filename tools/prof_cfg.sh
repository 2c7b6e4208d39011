#!/bin/bash
# on the GPU box: ncu --set full of ONE k_bounce launch of a BASELINE config rendered by tools/run_configs.py
# usage: tools/prof_cfg.sh <tag> <config-substring> [skip] [frac]
tag=$1; cfg=$2; skip=${3:-1}; frac=${4:-0.008}
ncu --set full --clock-control none --import-source on -k regex:k_bounce -s $skip -c 1 -o gpurun_out/${tag}_prof -f \
  python tools/run_configs.py --only $cfg --frac $frac --out gpurun_out/${tag}_cfg.jsonl > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_src.csv
