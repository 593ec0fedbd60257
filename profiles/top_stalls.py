#!/usr/bin/env python
"""Per-kernel top stall sites from `ncu --page source --csv` (SASS view): usage top_stalls.py rep.ncu-rep [regex] [N]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None and row:
        if cur["hdr"] is None:
            cur["hdr"] = row
        else:
            cur["rows"].append(row)
for b in blocks:
    if pat and not pat.search(b["name"]):
        continue
    h = {n: i for i, n in enumerate(b["hdr"])}
    si, ins = h["# Samples"], h["Instructions Executed"]
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print(f"== {b['name']}  samples={tot}")
    top = sorted(b["rows"], key=lambda r: -int(r[si] or 0))[:topn]
    for r in top:
        print(f"  {100*int(r[si] or 0)/max(tot,1):5.1f}%  exec={r[ins]:>8s}  {r[h['Source']].strip()[:110]}")
