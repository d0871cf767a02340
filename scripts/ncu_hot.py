#!/usr/bin/env python3
"""Hottest SASS instructions of one kernel in an .ncu-rep (stall samples and executed counts).
usage: scripts/ncu_hot.py <rep> <kernel regex> [top N]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first launch only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h0 = hdr_i[0]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
hdr = rows[h0]
body = [dict(zip(hdr, r)) for r in rows[h0 + 1:end] if len(r) == len(hdr)]
tot_s = sum(int(b["# Samples"] or 0) for b in body)
tot_i = sum(int(b["Instructions Executed"] or 0) for b in body)
print(rows[0][1][:100] if rows[0] else "", "samples", tot_s, "warp-instr", tot_i, "sass lines", len(body))
stall_cols = [c for c in hdr if c.startswith("stall_")]
for rank, b in enumerate(sorted(body, key=lambda b: -int(b["# Samples"] or 0))[:top]):
    st = sorted(((int(b[c] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{100.0 * int(b['# Samples']) / tot_s:5.1f}% smp  {100.0 * int(b['Instructions Executed']) / tot_i:4.1f}% ins  thr {b['Avg. Threads Executed']:>4}  "
          f"{b['Source'].strip()[:70]:70s} {' '.join(f'{c[6:]}={v}' for v, c in st if v)}")
