#!/usr/bin/env python
"""Top stall locations (SASS) of an ncu report with source info: python scripts/ncu_hot.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
for b in blocks[:1]:
    h = b["hdr"]
    si, ii = h.index("# Samples"), h.index("Source")
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print(b["name"], "total samples", tot)
    idx = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][si] or 0))[:n]
    for i in sorted(idx):
        r = b["rows"][i]
        print(f"{i:5d} {int(r[si]):7d} {100*int(r[si])/max(tot,1):5.1f}%  {r[ii].strip()[:110]}")
