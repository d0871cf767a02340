"""Aggregate an `ncu --page source --csv` dump by SASS opcode (executed warp instructions)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
        continue
    if cur is None:
        continue
    if r and r[0] == "Address":
        cur["hdr"] = r
        continue
    cur["rows"].append(r)
for k in kern[: int(sys.argv[2]) if len(sys.argv) > 2 else 2]:
    h = k["hdr"]
    ia, isrc = h.index("Instructions Executed"), h.index("Source")
    tot, byop = 0, collections.Counter()
    for r in k["rows"]:
        try:
            n = int(r[ia])
        except (ValueError, IndexError):
            continue
        tot += n
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        byop[m.group(2).split(".")[0] if m else "?"] += n
    print(k["name"][:60], "total warp-inst", tot, "static", len(k["rows"]))
    for op, n in byop.most_common(28):
        print("   %-10s %12d %5.1f%%" % (op, n, 100 * n / tot))
