#!/bin/bash
# One GPU round: parity tests, bench line, ICP launch list, one full ncu capture of k_knn.  $1 = tag.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 900 python -u bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo bench rc=$?
grep "cached iteration" gpurun_out/bench_$TAG.log; tail -1 gpurun_out/bench_$TAG.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'e2e ms',round(d['e2e']['ms_per_step'],3))
r=d['roofline']; print('knn us',round(r['us_per_launch'],1),'frac',round(r['frac'],3),'local us',round(r['local_regime']['us_per_launch'],1))
print('cpu',d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('faithful_4_threads'), 'clocks', d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_linearize|k_finalize|k_loc_comp" -c 130 --csv --log-file gpurun_out/launches_icp_$TAG.csv python -u bench.py --profile-icp --steps 2 > gpurun_out/ncu_icp.log 2>&1; echo ncu1 rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn -s 1 -c 1 -o gpurun_out/prof_knn_$TAG python -u bench.py --profile-knn --steps 3 > gpurun_out/ncu_knn.log 2>&1; echo ncu2 rc=$?
