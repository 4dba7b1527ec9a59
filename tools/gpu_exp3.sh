#!/bin/bash
run() { timeout 600 python -u bench.py --knn-only 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())['roofline']
print('$1: knn us',round(r['us_per_launch'],1),'min',round(r['us_min'],1),'local us',round(r['local_regime']['us_per_launch'],1))"; }
cp mimosa_b200/lib/libmimosa_b200.so /tmp/lib_keep.so
cp mimosa_b200/lib/libmimosa_b200_skipnb.so mimosa_b200/lib/libmimosa_b200.so
for nq in 4096 16384 131072; do MB_BENCH_NQ=$nq run "skipnb nq=$nq"; done
cp /tmp/lib_keep.so mimosa_b200/lib/libmimosa_b200.so
for nq in 4096; do MB_BENCH_NQ=$nq run "full nq=$nq"; done
