#!/usr/bin/env python
"""Hot SASS instructions (stall samples) of kernel number K in an .ncu-rep source page.

    python tools/ncu_sass.py rep.ncu-rep K [min_pct]
"""
import csv, subprocess, sys
rep, K = sys.argv[1], int(sys.argv[2])
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = secs[K]
hdr = sec["rows"][0]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in sec["rows"][1:] if len(r) > iex]
tot = sum(int(r[isamp]) for r in data)
print(sec["name"], "kernels:", len(secs), "samples:", tot, "instrs:", len(data))
for k, r in enumerate(data):
    s = int(r[isamp])
    if s > tot * minpct / 100 or any(t in r[ia] for t in ("UTCHMMA", "UBLKCP", "UTCBAR", "LDTM", "BAR.")):
        print(f"{k:5d} {r[ia].strip()[:96]:96s} {s:7d} {100*s/tot:5.1f}% ex={r[iex]}")
