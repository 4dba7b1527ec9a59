"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py file.csv [n_seq]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
per, seq = collections.defaultdict(list), []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].split("::")[-1]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    per[name].append(v)
    seq.append((name, v))
tot = sum(sum(v) for v in per.values())
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print("%-28s n=%4d total %9.1f us (%5.1f%%) mean %7.1f min %7.1f max %7.1f" % (k, len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v), min(v), max(v)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if n:
    print(" ".join("%s:%.0f" % (a[:6], b) for a, b in seq[:n]))
