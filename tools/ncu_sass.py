#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report: tools/ncu_sass.py <report.ncu-rep> [launch_index] [min_pct]
Prints every instruction with its share of warp-stall samples, executions and average active threads;
consecutive cold instructions are folded."""
import csv
import subprocess
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(which), "--launch-count", "1"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)
b = blocks[0]
h = {n: i for i, n in enumerate(b["hdr"])}
rows = b["rows"]
tot_s = sum(int(r[h["# Samples"]]) for r in rows) or 1
tot_i = sum(int(r[h["Instructions Executed"]]) for r in rows) or 1
tot_t = sum(int(r[h["Thread Instructions Executed"]]) for r in rows)
print("# %s\n# %d SASS instructions, %d samples, %d warp-instr executed, avg active threads %.2f" % (b["name"], len(rows), tot_s, tot_i, tot_t / tot_i))
fold_s = fold_i = fold_n = 0
for k, r in enumerate(rows):
    s, ie, te = int(r[h["# Samples"]]), int(r[h["Instructions Executed"]]), int(r[h["Thread Instructions Executed"]])
    sass = r[h["Source"]].strip()
    hot = 100.0 * s / tot_s >= min_pct or any(t in sass for t in ("LDG", "BRA", "STG", "ATOM", "VOTE", "SHFL", "EXIT", "LDS", "STS", "LDL", "STL", "WARPSYNC", "BSSY", "BSYNC"))
    if not hot:
        fold_s += s; fold_i += ie; fold_n += 1
        continue
    if fold_n:
        print("      ... %3d instr  samples %5.2f%%  exec %5.2f%%" % (fold_n, 100.0 * fold_s / tot_s, 100.0 * fold_i / tot_i))
        fold_s = fold_i = fold_n = 0
    stalls = sorted(((n[6:], int(r[i])) for n, i in h.items() if n.startswith("stall_") and "Not Issued" not in n and r[i] not in ("", "0")), key=lambda x: -x[1])[:2]
    print("%4d %5.2f%% exec %5.2f%% thr %4.1f  %-60s %s" % (k, 100.0 * s / tot_s, 100.0 * ie / tot_i, te / max(1, ie), sass[:60], " ".join("%s:%d" % x for x in stalls)))
if fold_n:
    print("      ... %3d instr  samples %5.2f%%  exec %5.2f%%" % (fold_n, 100.0 * fold_s / tot_s, 100.0 * fold_i / tot_i))
