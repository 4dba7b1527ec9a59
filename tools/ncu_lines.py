"""Per-CUDA-source-line summary of an ncu report: instructions executed and stall samples per line.
usage: python tools/ncu_lines.py report.ncu-rep [n_top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur_file, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ss, ln, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), 0, 1
        continue
    if hdr and r[0] not in ("", "Function Name") and r[0].isdigit():
        try:  # source text may contain quotes / commas: index the metric columns from the end
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ie - len(hdr)] or 0), int(r[ss - len(hdr)] or 0)))
        except ValueError:
            pass
tot_i = sum(x[3] for x in lines)
tot_s = sum(x[4] for x in lines)
print("total instr", tot_i, "samples", tot_s)
print("-- by instructions")
for f, n, s, i, sm in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%5.1f%% instr %5.1f%% samp  %s:%d  %s" % (100 * i / tot_i, 100 * sm / max(tot_s, 1), f, n, s[:110]))
print("-- by samples")
for f, n, s, i, sm in sorted(lines, key=lambda x: -x[4])[:25]:
    print("%5.1f%% instr %5.1f%% samp  %s:%d  %s" % (100 * i / tot_i, 100 * sm / max(tot_s, 1), f, n, s[:110]))
