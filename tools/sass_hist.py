"""Summarise an `ncu --page source --csv` SASS dump: weighted opcode histogram and hot contiguous regions.
usage: python tools/sass_hist.py file.csv n_units [kernel-substring]"""
import collections
import csv
import math
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
want = sys.argv[3] if len(sys.argv) > 3 else None
kernels, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "data": []}
        kernels.append(cur)
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if cur is not None and hdr and len(r) >= 8 and r[0].startswith("0x"):
        cur["data"].append(r)
ie, si, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
for kern in kernels:
    if want and want not in kern["name"]:
        continue
    data = kern["data"]
    tot = sum(int(r[ie]) for r in data)
    samp = sum(int(r[ss]) for r in data)
    print("==", kern["name"][:100], "| sass", len(data), "| exec", tot, "| per unit %.1f" % (tot / units), "| samples", samp)
    op = collections.Counter()
    sop = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += int(r[ie])
        sop[o] += int(r[ss])
    for o, c in op.most_common(28):
        print("  %-10s %6.1f%%  %7.1f/unit   stall-samples %5.1f%%" % (o, 100 * c / tot, c / units, 100 * sop[o] / max(samp, 1)))
    print("  -- regions (contiguous, similar exec count)")
    regs, lv, start, acc, cnt, sacc = [], None, 0, 0, 0, 0
    for i, r in enumerate(data):
        e = int(r[ie])
        l = int(math.log2(e + 1) * 2)
        if lv is None:
            lv, start = l, i
        if l != lv:
            regs.append((start, i - 1, acc, cnt, sacc))
            lv, start, acc, cnt, sacc = l, i, 0, 0, 0
        acc += e
        cnt += 1
        sacc += int(r[ss])
    regs.append((start, len(data) - 1, acc, cnt, sacc))
    for s, e, a, c, sa in regs:
        if a / tot > 0.015 or sa / max(samp, 1) > 0.02:
            print("  [%4d-%4d] n=%4d exec/instr=%9d share %5.1f%% samples %5.1f%% | %s" % (
                s, e, c, a // c, 100 * a / tot, 100 * sa / max(samp, 1), data[s][si].strip()[:50]))
