#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/ncu_launch_summary.py list.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("b200::", "")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[13]]
    tot[name] += float(r[14]) * scale
    cnt[name] += 1
total = sum(tot.values())
print("%-60s %6s %12s %7s" % ("kernel", "calls", "total_ms", "share"))
for name in sorted(tot, key=lambda n: -tot[n]):
    print("%-60s %6d %12.3f %6.1f%%" % (name[:60], cnt[name], tot[name], 100 * tot[name] / total))
