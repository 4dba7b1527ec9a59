#!/bin/bash
# brief summary of an ncu report: tools/ncu_brief.sh file.ncu-rep n_units
ncu -i $1 --page details --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); H=rows[0]; idx={h:i for i,h in enumerate(H)}
want=['Duration','Registers Per Thread','Achieved Occupancy','Theoretical Occ','DRAM Throughput','L2 Cache Throughput','L1/TEX Hit Rate','L2 Hit Rate','Executed Ipc Active','Issue Slots Busy','No Eligible','Warp Cycles Per Issued','Avg. Active Threads','Waves Per SM','Executed Instructions','L1/TEX Cache Throughput','Local']
seen=set()
for r in rows[1:]:
    n=r[idx['Metric Name']]
    if any(w.lower() in n.lower() for w in want) and n not in seen: seen.add(n); print(n.ljust(48), r[idx['Metric Value']], r[idx['Metric Unit']])
"
ncu -i $1 --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); H=rows[0]
for name in H:
    if name in ('dram__bytes_read.sum','dram__bytes_write.sum') or ('warp_issue_stalled' in name and name.endswith('per_warp_active.pct')):
        v=rows[2][H.index(name)]
        try:
            if 'stalled' in name and float(v)<3: continue
        except: pass
        print(name, v, rows[1][H.index(name)])
"
ncu -i $1 --page source --csv 2>/dev/null > /tmp/_src.csv; python tools/sass_hist.py /tmp/_src.csv $2 | sed -n '1,14p;30,60p'
