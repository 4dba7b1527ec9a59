#!/bin/bash
# gpurun_out/prof_r2_*.ncu-rep + launch list (tools/gpu_final.sh) -> profiles/r2_*_summary.txt, roofline_traffic.json, SASS
set -e
SHA=$(git rev-parse --short HEAD)
(echo "# ncu --set full --clock-control none --import-source on, k_icp_loop<5, 19>: ONE launch = the whole 20-iteration ICP loop of a bench scan (131 072 points, 10.1 M-point map, L2 flushed before the scan); library built from commit $SHA"
 bash tools/ncu_brief.sh gpurun_out/prof_r2_icp_loop.ncu-rep 4096 2>&1 | head -60
 echo; echo "# warp stall samples by reason"
 ncu -i gpurun_out/prof_r2_icp_loop.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); H=rows[0]
out=[]
for name in H:
    if 'pcsamp_warps_issue_stalled' in name and 'not_issued' not in name:
        try: out.append((float(rows[2][H.index(name)]), name))
        except: pass
tot=sum(v for v,_ in out)
for v,n in sorted(out,reverse=True)[:10]: print('%6.1f%%  %s'%(100*v/tot, n.replace('smsp__pcsamp_warps_issue_stalled_','')))
"
 echo; echo "# hottest source lines (tools/ncu_lines.py)"
 python tools/ncu_lines.py gpurun_out/prof_r2_icp_loop.ncu-rep 25 2>&1 | head -60) > profiles/r2_k_icp_loop_ncu_summary.txt
(echo "# ncu --set full --clock-control none --import-source on, k_knn<5>: 131 072 spread queries vs the 10.1 M-point map, L2 flushed (256 MiB write) before the launch; library built from commit $SHA"
 bash tools/ncu_brief.sh gpurun_out/prof_r2_knn.ncu-rep 131072 2>&1 | head -60) > profiles/r2_k_knn_spread_ncu_summary.txt
python - <<PY
import csv, subprocess, json
out = subprocess.run(["ncu","-i","gpurun_out/prof_r2_knn.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); H=rows[0]
def val(name):
    v=rows[2][H.index(name)]; u=rows[1][H.index(name)]
    return int(round(float(v.replace(",",""))*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[u]))
r,w=val("dram__bytes_read.sum"),val("dram__bytes_write.sum")
d={"k_knn_dram_bytes_per_launch": r+w, "dram_bytes_read": r, "dram_bytes_write": w,
   "source": "ncu --set full --clock-control none, k_knn<5>, 131072 spread queries vs 10.1M-pt map, L2 flushed (256 MiB write) before the launch; round 2, library built from commit $SHA (tools/gpu_final.sh -> profiles/r2_k_knn_spread_ncu_summary.txt)"}
json.dump(d, open("profiles/roofline_traffic.json","w"), indent=1)
print("traffic", r+w)
PY
cp gpurun_out/r2_launches_icp_3steps.csv profiles/r2_launches_icp_3steps.csv
(echo "# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_icp_loop|k_group_sort|k_flush|k_pack_src', python bench.py --profile-icp --steps 3 (three 20-iteration scans; per-launch times are serialised and cold: shares, not absolutes)"
 python tools/launch_summary.py profiles/r2_launches_icp_3steps.csv 12) > profiles/r2_launches_icp_summary.txt
bash tools/dump_sass.sh
