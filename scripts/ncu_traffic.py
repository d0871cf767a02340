#!/usr/bin/env python3
"""Write profiles/r2_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the volume
kernels from one `ncu --set full` capture of the bench workload.  bench.py prints these as `roofline.traffic`.
usage: scripts/ncu_traffic.py <file.ncu-rep> <size> [out.json]"""
import csv, io, json, subprocess, sys

NAMES = {"stats_fast_kernel": "nb200_hessian_stats_fast", "frangi_fast_kernel": "nb200_frangi_fast",
         "gauss_z_vec": "nb200_gauss_axis", "gauss_yx_tile": "nb200_gauss_yx", "opening_march_kernel": "nb200_finalize_opening"}
rep, size = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/r2_traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
acc = {}
for r in rows[2:]:
    for key, name in NAMES.items():
        if key in r[ik]:
            b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            acc.setdefault(name, []).append(b)
res = {"size": size, "source": rep.split("/")[-1], "how": "ncu --set full --clock-control none, mean over the captured launches",
       "kernels": {k: sum(v) / len(v) for k, v in acc.items()}, "launches": {k: len(v) for k, v in acc.items()}}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
