#!/usr/bin/env python
"""stdin: `ncu -i rep --page source --csv` -> the 25 source lines with the most warp-stall samples (needs -lineinfo)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
if not rows:
    sys.exit(0)
hdr = None
for i, r in enumerate(rows):
    if any(c.strip() in ("Source", "# Samples", "Warp Stall Sampling (All Samples)") for c in r):
        hdr, body = r, rows[i + 1:]
        break
if hdr is None:
    print("no source table"); sys.exit(0)
col = {h.strip(): i for i, h in enumerate(hdr)}
ks = [k for k in ("Warp Stall Sampling (All Samples)", "# Samples", "Samples") if k in col]
src = col.get("Source")
if not ks or src is None:
    print("columns:", list(col)[:20]); sys.exit(0)
k = col[ks[0]]
agg = {}
for r in body:
    try:
        v = float(r[k].replace(",", ""))
    except (ValueError, IndexError):
        continue
    agg[r[src].strip()] = agg.get(r[src].strip(), 0.0) + v
tot = sum(agg.values()) or 1.0
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{100 * v / tot:6.2f}%  {s[:150]}")
