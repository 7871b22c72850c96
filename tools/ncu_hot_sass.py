#!/usr/bin/env python
"""Top stall sites of one kernel instance from an ncu source page dump.
usage: ncu -i rep --page source --csv --print-source sass | python tools/ncu_hot_sass.py <instance index> [top N]"""
import csv
import sys

want = int(sys.argv[1]) if len(sys.argv) > 1 else 0
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(sys.stdin))
inst, blocks, cur = -1, [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
b = blocks[want]
h = b["hdr"]
iS, iSamp, iN, iEx = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[iSamp]) for r in b["rows"])
print("kernel instance %d: %d instructions, %d stall samples" % (want, len(b["rows"]), tot))
agg = {}
for r in b["rows"]:
    for i in stalls:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
order = sorted(range(len(b["rows"])), key=lambda j: -int(b["rows"][j][iSamp]))[:top]
for j in sorted(order):
    r = b["rows"][j]
    why = sorted(((int(r[i]), h[i][6:]) for i in stalls if int(r[i])), reverse=True)[:3]
    print("%5d %6.2f%% ex=%-8s %-70s %s" % (j, 100.0 * int(r[iSamp]) / max(tot, 1), r[iEx], r[iS].strip()[:70], why))
