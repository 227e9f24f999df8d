#!/usr/bin/env python
"""Per-source-line stall summary of an `ncu --page source --print-source cuda,sass --csv` export.
usage: ncu_source_summary.py export.csv [top_n]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows, header, fname = [], None, None
per_line = []
for rec in csv.reader(open(path, newline="")):
    if not rec:
        continue
    if rec[0] == "File Path":
        fname = rec[1].split("/")[-1]
        continue
    if rec[0] == "Function Name":
        continue
    if rec[0] == "Line No":
        header = rec
        continue
    if header is None or rec[0] == "":
        continue  # SASS rows: already summed into their source line
    d = dict(zip(header, rec))
    def num(k):
        try:
            return float(d.get(k, "0") or 0)
        except ValueError:
            return 0.0
    stalls = {k[6:]: num(k) for k in header if k.startswith("stall_") and "Not Issued" not in k}
    per_line.append((fname, int(rec[0]), rec[1].strip()[:70], num("# Samples"), num("Instructions Executed"), stalls))
total = sum(p[3] for p in per_line) or 1.0
print(f"total samples {total:.0f}")
agg = defaultdict(float)
for p in per_line:
    for k, v in p[5].items():
        agg[k] += v
print("stall reasons overall:", ", ".join(f"{k} {100*v/total:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for f, ln, src, smp, ins, st in sorted(per_line, key=lambda p: -p[3])[:top]:
    main = ", ".join(f"{k} {100*v/max(smp,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*smp/total:5.1f}%  {f}:{ln:<4d} inst {ins:10.0f}  [{main}]  {src}")
