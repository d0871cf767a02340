#!/usr/bin/env python3
"""Opcode histogram of the largest loop (longest backward-branch span) of one kernel in an object file.
usage: scripts/sass_loop.py <file.o> <substring of the mangled kernel name> [--dump]"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, txt in ins:
        m = re.search(r"BRA\S*\s+(?:[!\w]+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?", txt)
        m2 = re.search(r"BRA.*0x([0-9a-f]+)", txt)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    print(name[:120])
    print("instructions:", len(ins), "loop:", best and (hex(best[0]), hex(best[1])))
    if best:
        body = [t for a, t in ins if best[0] <= a <= best[1]]
        h = collections.Counter()
        for t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            h[t.split()[0].split(".")[0]] += 1
        print("loop body:", len(body))
        print(", ".join(f"{k}:{v}" for k, v in h.most_common()))
        if "--dump" in sys.argv:
            for a, t in ins:
                if best[0] <= a <= best[1]:
                    print(f"{a:05x}  {t}")
