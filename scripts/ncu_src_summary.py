#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: instructions executed and stall samples per
region of the kernel (regions = runs of SASS lines between the markers given on the command line) and the top lines.

    ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv ; python scripts/ncu_src_summary.py x.csv [top_n]
"""
import csv
import sys

path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {name: i for i, name in enumerate(hdr)}
data = rows[hdr_i + 1 :]
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]


def num(r, name):
    try:
        return float(r[col[name]])
    except (ValueError, IndexError):
        return 0.0


tot_inst = sum(num(r, "Instructions Executed") for r in data)
tot_samp = sum(num(r, "# Samples") for r in data)
print(f"lines={len(data)} instructions_executed={tot_inst:.3e} samples={tot_samp:.0f}")
print("stall totals:", ", ".join(f"{n[6:]}={sum(num(r, n) for r in data):.0f}" for n in stall_cols if sum(num(r, n) for r in data) > 0))
print(f"\n# top {top_n} lines by samples: idx, samples, inst_executed, avg_threads, top stalls, sass")
order = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:top_n]
for i in sorted(order):
    r = data[i]
    st = sorted(((num(r, n), n[6:]) for n in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {num(r,'# Samples'):7.0f} {num(r,'Instructions Executed'):10.0f} {num(r,'Avg. Threads Executed'):5.1f}  "
          f"{st[0][1]}={st[0][0]:.0f},{st[1][1]}={st[1][0]:.0f}  {r[col['Source']].strip()}")
# coarse profile: cumulative instructions / samples per 100 lines
print("\n# per 100 SASS lines: first idx, inst_executed share, samples share, first instruction")
for b in range(0, len(data), 100):
    blk = data[b : b + 100]
    ie = sum(num(r, "Instructions Executed") for r in blk)
    sm = sum(num(r, "# Samples") for r in blk)
    print(f"{b:5d} inst={100*ie/tot_inst:5.1f}% samples={100*sm/tot_samp:5.1f}%  {blk[0][col['Source']].strip()[:60]}")
