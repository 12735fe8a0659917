#!/usr/bin/env python
"""Static SASS census of libalphafive.so: per kernel, how many tcgen05 / TMEM / bulk-copy / barrier instructions it
contains (cuobjdump -sass; runs without a GPU).   python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "alphafive_b200", "libalphafive.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "SYNCS", "ELECT", "UCGABAR", "LDGSTS", "MUFU.EX2", "HMMA", "FFMA"]
arch, name, per = None, None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        per[name] = collections.Counter(arch=arch)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        per[name]["instr"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[name][k] += 1
print(f"{os.path.basename(lib)}: cuobjdump -sass, instruction counts per kernel (static)")
print(f"{'kernel':44s} {'arch':8s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
tot = collections.Counter()
for n, c in per.items():
    short = re.sub(r"^a5::", "", n)[:44]
    print(f"{short:44s} {c['arch']:8s} {c['instr']:6d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
    tot.update({k: c[k] for k in KEYS})
print(f"{'total':44s} {'':8s} {'':6s} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA engine, 1-D), "
      "UTMALDG = tensor-map TMA (not used: DESIGN 3.2), SYNCS = mbarrier, UCGABAR = cluster barrier, LDGSTS = cp.async")
