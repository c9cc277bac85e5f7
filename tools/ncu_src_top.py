#!/usr/bin/env python
"""Top stall locations of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view): prints the instructions with the
most warp-stall samples, their dominant stall reason and executed count."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# the file may hold several kernels; split at "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if cur is not None: cur["rows"].append(r)
for b in blocks:
    hdr = b["rows"][0]; body = b["rows"][1:]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp] or 0) for r in body)
    totex = sum(int(r[iex] or 0) for r in body)
    print(f"== {b['name']}  samples {tot}  warp-instructions {totex}  SASS lines {len(body)}")
    agg = {}
    for r in body:
        for i, h in stalls:
            agg[h] = agg.get(h, 0) + int(r[i] or 0)
    print("   stall mix:", ", ".join(f"{h[6:]} {100*v/max(tot,1):.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(body)), key=lambda k: -int(body[k][isamp] or 0))[:top]
    for k in sorted(order):
        r = body[k]
        s = int(r[isamp] or 0)
        dom = max(stalls, key=lambda ih: int(r[ih[0]] or 0))
        print(f"   #{k:5d} {100*s/max(tot,1):5.2f}%  ex {int(r[iex] or 0):9d}  {dom[1][6:]:12s} {r[isrc].strip()[:90]}")
