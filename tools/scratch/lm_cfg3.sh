#!/bin/bash
# per-launch metrics of the first k_bounce_bvh launches on config 3 for a library variant
for v in "$@"; do
  if [ "$v" = base ]; then lib=""; else lib="build/libpt_$v.so"; fi
  PT_B200_LIB=$lib ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum \
    --clock-control none -k regex:k_bounce_bvh -c 3 --csv --log-file gpurun_out/lm_$v.csv \
    python tools/run_configs.py --only config3 --frac 0.008 --out /tmp/x.jsonl > /dev/null 2>&1
  echo "== $v"
  python - gpurun_out/lm_$v.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r[0], {})[r[-3]] = r[-1]
for i, d in by.items():
    print("  ".join("%s=%s" % (k.split("__")[-1][:34], v) for k, v in d.items()))
PY
done
