#!/usr/bin/env python
"""Dynamic instruction mix per kernel from `ncu --page source --csv` (needs --import-source / SourceCounters):
usage instruction_mix.py rep.ncu-rep [kernel-regex] [units-per-launch: e.g. unique triples]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
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
seen = set()
for b in blocks:
    if (pat and not pat.search(b["name"])) or b["name"] in seen:
        continue
    seen.add(b["name"])
    h = {n: i for i, n in enumerate(b["hdr"])}
    tot, ops = 0, collections.Counter()
    for r in b["rows"]:
        n = int(r[h["Instructions Executed"]] or 0)
        tot += n
        src = r[h["Source"]].strip().split()
        op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
        ops[op.split(".")[0]] += n
    print(f"== {b['name'][:90]}\n   warp instructions executed: {tot:,}")
    print("   " + ", ".join(f"{k} {100*v/max(tot,1):.1f}%" for k, v in ops.most_common(14)))
