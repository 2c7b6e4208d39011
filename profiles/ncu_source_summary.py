#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass`: executed warp-instructions and stall samples per
CUDA source line and per SASS opcode, for one kernel.
usage: python profiles/ncu_source_summary.py src.csv 'k_bounce<(bool)0, (bool)0>' [topN]"""
import csv, sys, collections, re


def num(s):
    try:
        return int(float(s.replace(",", "")))
    except ValueError:
        return 0

path, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}; blocks.append(cur)
    elif r and r[0] == "Function Name" and cur is not None:
        cur["fn"] = r[1]
    elif r and r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r:
        cur["rows"].append(r)
done = set()
for b in blocks:
    if kern not in b.get("fn", "") or (b["fn"], b["file"]) in done:
        continue
    H = b["hdr"]
    # the SASS view: rows with an Address
    iaddr, iline = H.index("Address"), H.index("Line No")
    isrc = [i for i, h in enumerate(H) if h == "Source"]
    iex, ith, ismp = H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("# Samples")
    sass = [r for r in b["rows"] if len(r) > iaddr and r[iaddr].startswith("0x")]
    if not sass:
        continue
    done.add((b["fn"], b["file"]))
    tot = sum(num(r[iex]) for r in sass); tott = sum(num(r[ith]) for r in sass)
    print("== %s [%s]: %d SASS instr, %d warp-inst executed, %.1f avg active threads" % (b["fn"], b["file"].split("/")[-1], len(sass), tot, tott / max(tot, 1)))
    ops = collections.Counter(); smp = collections.Counter()
    for r in sass:
        op = r[isrc[1]].split()
        op = [o for o in op if not o.startswith("@")]
        name = op[0].split(".")[0] if op else "?"
        ops[name] += num(r[iex]); smp[name] += num(r[ismp])
    print("  opcode mix (warp-inst %, stall-sample %):")
    ts = sum(smp.values())
    for k, v in ops.most_common(top):
        print("    %-10s %5.1f%%  %5.1f%%" % (k, 100.0 * v / tot, 100.0 * smp[k] / max(ts, 1)))
# per CUDA line
seen = set()
for b in blocks:
    if kern not in b.get("fn", "") or b["file"] in seen:
        continue
    seen.add(b["file"])
    H = b["hdr"]; iaddr = H.index("Address")
    src = [r for r in b["rows"] if len(r) > iaddr and r[0].strip().isdigit()]
    if not src:
        continue
    iex, ismp, iline = H.index("Instructions Executed"), H.index("# Samples"), H.index("Line No")
    isrc = [i for i, h in enumerate(H) if h == "Source"][0]
    tot = sum(num(r[iex]) for r in src); ts = sum(num(r[ismp]) for r in src)
    if tot == 0:
        continue
    print("== per source line, file %s (%.1f M warp-inst)" % (b["file"].split("/")[-1], tot / 1e6))
    for r in sorted(src, key=lambda r: -num(r[iex]))[:top]:
        print("    %5s %5.1f%% %5.1f%%  %s" % (r[iline], 100.0 * num(r[iex]) / tot, 100.0 * num(r[ismp]) / max(ts, 1), r[isrc].strip()[:110]))
