#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump (SASS view) into phases delimited by marker instructions.
Usage: ncu_phases.py dump.csv [marker-regex]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
marker = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"UTCHMMA|UCGABAR|BAR\.SYNC|UTCBAR|STG|LDG|UBLKCP")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]]) for r in rows[2:])
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in rows[2:])
print("total samples", tot, "instructions", tot_inst)
seg = {"samples": 0, "inst": 0, "st": {s: 0 for s in stalls}, "n": 0, "wf": 0, "wf_ideal": 0}
def flush(label):
    if seg["n"] == 0:
        return
    top = sorted(seg["st"].items(), key=lambda kv: -kv[1])[:4]
    print(f"{seg['samples']:7d} ({100*seg['samples']/tot:5.1f}%) inst {seg['inst']:9d} ({100*seg['inst']/tot_inst:5.1f}%) smem-wf {seg['wf']:9d}/{seg['wf_ideal']:9d}  "
          + " ".join(f"{k[6:]}={v}" for k, v in top if v) + f"   -> {label}")
    seg.update({"samples": 0, "inst": 0, "st": {s: 0 for s in stalls}, "n": 0, "wf": 0, "wf_ideal": 0})
for r in rows[2:]:
    sass = r[col["Source"]].strip()
    seg["samples"] += int(r[col["# Samples"]])
    seg["inst"] += int(r[col["Instructions Executed"]])
    seg["wf"] += int(r[col["L1 Wavefronts Shared"]] or 0)
    seg["wf_ideal"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    seg["n"] += 1
    for s in stalls:
        seg["st"][s] += int(r[col[s]] or 0)
    if marker.search(sass):
        flush(sass[:60])
flush("end")
