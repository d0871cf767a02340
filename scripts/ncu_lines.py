"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
executed warp instructions, thread instructions and stall samples, sorted by instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fpath, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0].isdigit():
        ie = hdr.index("Instructions Executed")
        te = hdr.index("Thread Instructions Executed")
        ss = hdr.index("# Samples")
        try:
            out.append((int(r[ie]), int(r[te]), int(r[ss]), fpath, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot_i = sum(o[0] for o in out)
tot_s = sum(o[2] for o in out)
print("total warp-inst %d  samples %d" % (tot_i, tot_s))
byfile = {}
for o in out:
    a = byfile.setdefault(o[3], [0, 0])
    a[0] += o[0]
    a[1] += o[2]
for f, (i, s) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-24s inst %5.1f%%  samples %5.1f%%" % (f, 100 * i / tot_i, 100 * s / max(1, tot_s)))
for o in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.1f%% i %5.1f%% s  thr/inst %4.1f  %s:%d  %s" % (100 * o[0] / tot_i, 100 * o[2] / max(1, tot_s), o[1] / max(1, o[0]), o[3], o[4], o[5][:90]))
