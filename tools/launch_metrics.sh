#!/bin/bash
# on the GPU box: per-launch duration / DRAM bytes / instructions / issue utilisation of the first N k_* launches of a
# one-wavefront render.  usage: tools/launch_metrics.sh <tag> [n_launches] [extra bench args]
tag=$1; n=${2:-16}; shift; shift
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
  --clock-control none -k regex:k_ -c $n --csv --log-file gpurun_out/${tag}_lm.csv \
  python bench.py --steps 1 --warmup 0 --spp 16 --wf-spp 16 --no-cpu-baseline "$@" > gpurun_out/${tag}_lm.log 2>&1
python - "$tag" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open("gpurun_out/%s_lm.csv" % sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r[0], {"k": r[4]})[r[-3]] = r[-1]
for i, d in by.items():
    g = lambda k: float(d.get(k, "0").replace(",", ""))
    print("%-42s %8.1f us  rd %7.1f MB  wr %7.1f MB  inst %7.1f M  issue %5.1f%%  thr/inst %5.2f" % (
        d["k"].replace("void ptd::", "")[:42], g("gpu__time_duration.sum") / 1e3, g("dram__bytes_read.sum") / 1e6,
        g("dram__bytes_write.sum") / 1e6, g("smsp__inst_executed.sum") / 1e6,
        g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("smsp__thread_inst_executed_per_inst_executed.ratio")))
PY
