#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of a depth-1 k_bounce launch (tools/prof.sh):
DRAM bytes / algorithmic bytes of that launch, and the pipe utilisations bench.py quotes.
usage: python tools/update_traffic.py gpurun_out/<tag>_raw.csv profiles/<summary>.txt
The captured launch is the depth-1 launch of the timed 32-spp wavefront of `bench.py --spp 32` (seed 565): its live
counts are deterministic -- 18 015 703 paths in, 12 008 108 out (bench JSON `live_per_depth` x 32 / 5000)."""
import csv, json, os, sys

raw, summary = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
H, V = rows[0], rows[2]
g = lambda k: float(V[H.index(k)].replace(",", ""))
unit = lambda k: rows[1][H.index(k)]
def byt(k):
    v, u = g(k), unit(k)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
def us(k):
    v, u = g(k), unit(k)
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[u]
paths_in, paths_out = 18015703, 12008108
alg = 48.0 * paths_in + 48.0 * paths_out
rd, wr = byt("dram__bytes_read.sum"), byt("dram__bytes_write.sum")
inst = g("smsp__inst_executed.sum")
out = {
    "kernel": V[H.index("Kernel Name")],
    "source": "%s (ncu --set full, one depth-1 launch of a 32-spp wavefront)" % summary,
    "profiled_launch": {"paths_in": paths_in, "paths_out": paths_out, "dram_bytes_read": rd, "dram_bytes_write": wr,
                        "dram_bytes": rd + wr, "algorithmic_bytes": alg, "duration_us": us("gpu__time_duration.sum")},
    "dram_over_algorithmic": (rd + wr) / alg,
    "pipes_pct_of_peak": {
        "issue_slots": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fma": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "alu": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "xu": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "lsu": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "dram": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warp_instructions_per_warp_segment": round(inst / (paths_in / 32.0)),
        "threads_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
    },
}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
if os.path.exists(path):  # keep what tools/update_inst.py measured (instructions per segment of a whole wavefront)
    old = json.load(open(path))
    for k in ("warp_inst_per_segment", "threads_per_instruction_wavefront", "inst_source", "sm_count"):
        if k in old:
            out[k] = old[k]
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
