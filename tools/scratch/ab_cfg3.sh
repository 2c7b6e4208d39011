#!/bin/bash
# A/B of library variants on config 3: tools/scratch/ab_cfg3.sh base mb4 VAR=1:r4 ...   (ENV=val:variant sets an env var)
for spec in "$@"; do
  envs=""; v=$spec
  if [[ "$spec" == *:* ]]; then envs="${spec%%:*}"; v="${spec##*:}"; fi
  if [ "$v" = base ]; then lib=""; else lib="build/libpt_$v.so"; fi
  r=$(env $envs PT_B200_LIB=$lib timeout 300 python tools/run_configs.py --only config3 --frac 0.06 --out /tmp/ab.jsonl 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.1f Mseg/s  %.2f ms  segs %d" % (d["Mseg_per_s"], d["render_ms"], d["segments"]))')
  echo "$spec: $r"
done
