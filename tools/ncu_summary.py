#!/usr/bin/env python
"""Summarise `ncu --set full` reports into profiles/: one CSV row per captured kernel launch with the metrics the
roofline discussion uses, and the per-launch DRAM traffic table bench.py reports as `roofline.traffic`.
Usage: ncu_summary.py OUT_PREFIX report1.ncu-rep [report2.ncu-rep ...]   (needs the `ncu` CLI; no GPU)"""
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
UNIT_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

out_prefix, reports = sys.argv[1], sys.argv[2:]
rows_out, traffic = [], {}
for rep in reports:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[col["Kernel Name"]])
        short = re.match(r"(?:omk::)?(\w+)", name).group(1)
        rec = {"report": os.path.basename(rep), "kernel": name[:90]}
        for m in METRICS:
            if m in col:
                rec[m + " [" + units[col[m]] + "]"] = r[col[m]]
        rows_out.append(rec)
        rd = float(r[col["dram__bytes_read.sum"]]) * UNIT_BYTES[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * UNIT_BYTES[units[col["dram__bytes_write.sum"]]]
        key = short + ("_heads" if ", 128, 1" in name or ", (int)128, (bool)1" in name else "_fc1" if "<512" in name or "<(int)512" in name else "")
        traffic.setdefault(key, []).append(rd + wr)
keys = []
for rec in rows_out:
    for k in rec:
        if k not in keys:
            keys.append(k)
with open(out_prefix + "_summary.csv", "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=keys)
    w.writeheader()
    w.writerows(rows_out)
json.dump({k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v)} for k, v in traffic.items()},
          open(os.path.join(os.path.dirname(out_prefix) or ".", (re.match(r"(r\d+)_", os.path.basename(out_prefix)) or [None, "r00"])[1] + "_ncu_traffic.json"), "w"), indent=1)
print(open(out_prefix + "_summary.csv").read()[:3000])
