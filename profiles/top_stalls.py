#!/usr/bin/env python
"""Per-kernel top stall sites from `ncu --page source --csv` (SASS view), with the dominant stall reason and the two
preceding instructions for context: usage top_stalls.py rep.ncu-rep [kernel-regex] [N]"""
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
    reasons = [n for n in b["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print(f"== {b['name']}  samples={tot}")
    agg = {n: sum(int(r[h[n]] or 0) for r in b["rows"]) for n in reasons}
    print("   by reason: " + ", ".join(f"{n[6:]} {100*v/max(tot,1):.0f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][si] or 0))[:topn]
    for i in order:
        r = b["rows"][i]
        top_reason = max(reasons, key=lambda n: int(r[h[n]] or 0))
        ctx = " <- ".join(b["rows"][k][h["Source"]].strip()[:48] for k in range(i - 1, max(i - 3, -1), -1))
        print(f"  {100*int(r[si] or 0)/max(tot,1):5.1f}% {top_reason[6:]:>9s} exec={r[ins]:>7s}  {r[h['Source']].strip()[:60]:60s} | {ctx}")
