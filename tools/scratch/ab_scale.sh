#!/bin/bash
for sc in 1.0 0.5 0.25 0.0625; do
  r=$(PT_B200_LIB=$1 timeout 300 python tools/run_configs.py --only config3 --frac 0.06 --filter-scale $sc --out /tmp/ab.jsonl 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.1f Mseg/s  %.2f ms  segs %d fallback %.4f" % (d["Mseg_per_s"], d["render_ms"], d["segments"], d["fallback_fraction"]))')
  echo "scale $sc: $r"
done
