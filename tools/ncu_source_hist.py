#!/usr/bin/env python
"""Instruction-level view of one kernel from an `ncu --set full --import-source on` report (no GPU needed): executed warp
instructions per opcode, per-opcode stall samples, and the instructions that collected the most stall samples.
This is the analysis that found the per-MMA elect / R2UR waterfall code and the barrier waits in k_tower16.
Usage: ncu_source_hist.py REPORT.ncu-rep KERNEL_REGEX [TOP_N]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
print(rows[0][1][:120] if rows and len(rows[0]) > 1 else "")
h = rows[hdr_i]
data = []
for r in rows[hdr_i + 1:]:  # the first matching launch only (a report with several launches repeats the header block)
    if len(r) <= 5:
        continue
    if not r[0].startswith("0x"):
        break
    data.append(r)
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
warps = int(data[0][iE])  # the first instruction is executed once by every warp
tot = sum(int(r[iE]) for r in data)
tot_s = sum(int(r[iN]) for r in data) or 1
print(f"warps {warps}  executed warp instructions {tot}  per warp {tot / warps:.0f}  stall samples {tot_s}")
op, ops = collections.Counter(), collections.Counter()
for r in data:
    s = re.sub(r"^@!?U?P\w+\s+", "", r[iS].strip())
    o = s.split()[0].split(".")[0]
    op[o] += int(r[iE])
    ops[o] += int(r[iN])
print("\nopcode          per warp   share   stall samples")
for o, c in op.most_common(top_n):
    print(f"{o:14s} {c / warps:9.1f} {100 * c / tot:6.1f}%   {100 * ops[o] / tot_s:6.1f}%")
print("\ninstructions with the most stall samples (index, executed per warp, samples, SASS)")
order = sorted(range(len(data)), key=lambda i: -int(data[i][iN]))[:top_n]
for i in order:
    print(f"{i:6d} {int(data[i][iE]) / warps:8.1f} {int(data[i][iN]):6d}  {data[i][iS].strip()[:100]}")
