#!/usr/bin/env python
"""profiles/traffic.json: warp instructions per path segment of one whole wavefront, from the launch list that
tools/launch_metrics.sh wrote (ncu smsp__inst_executed.sum of every bounce-kernel launch of a 16-spp wavefront, and the
bench JSON of that same run for its segment count).  bench.py turns it into the issue-slot roofline.
usage: python tools/update_inst.py <tag>      (reads gpurun_out/<tag>_lm.csv and gpurun_out/<tag>_lm.log)"""
import csv, json, os, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", tag + "_lm.csv"))) if len(r) > 10 and r[0] != "ID"]
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r[0], {"k": r[4]})[r[-3]] = float(r[-1].replace(",", ""))
inst, launches, lanes_w = 0.0, 0, 0.0
for d in by.values():
    if "k_accum_counts" in d["k"] and launches:
        break  # one wavefront only: its segment count is what the log line states
    if "k_bounce" in d["k"]:
        inst += d["smsp__inst_executed.sum"]
        lanes_w += d["smsp__inst_executed.sum"] * d["smsp__thread_inst_executed_per_inst_executed.ratio"]
        launches += 1
line = [ln for ln in open(os.path.join(ROOT, "gpurun_out", tag + "_lm.log")) if ln.startswith("{")][-1]
segs = float(json.loads(line)["segments_per_step"])
path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(path))
tj["warp_inst_per_segment"] = inst / segs
tj["threads_per_instruction_wavefront"] = lanes_w / inst
tj["inst_source"] = "ncu launch list of one 16-spp wavefront (%d bounce launches, %.0f M warp instructions, %.0f segments), tools/launch_metrics.sh %s" % (launches, inst / 1e6, segs, tag)
tj["sm_count"] = 148
json.dump(tj, open(path, "w"), indent=1)
print("warp_inst_per_segment %.3f over %d launches, %.2f threads per instruction" % (inst / segs, launches, lanes_w / inst))
