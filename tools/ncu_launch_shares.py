#!/usr/bin/env python
"""Shares of the serialised kernel time in an `ncu --metrics gpu__time_duration.sum` launch list (B200_PROFILING.md
recipe): compare SHARES with bench.py's roofline `share_of_step`, not absolutes (ncu times are cold-cache, serialised).
Usage: ncu_launch_shares.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, tot, cnt = None, collections.Counter(), collections.Counter()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(d["Metric Unit"], 1)
    m = re.match(r"(?:void )?(?:omk::)?(\w+)(<[^>]*>)?", d["Kernel Name"])
    key = m.group(1) + (m.group(2) or "")
    tot[key] += v
    cnt[key] += 1
total = sum(tot.values())
print(f"{sum(cnt.values())} launches, {total / 1e3:.2f} ms of serialised kernel time")
for k, v in tot.most_common():
    print(f"{k[:48]:48s} {cnt[k]:6d} launches {v:10.1f} us {100 * v / total:6.1f} %   avg {v / cnt[k]:8.1f} us")
