#!/bin/bash
# development experiments: $1 = lib variant dir suffix list
run() { timeout 600 python -u bench.py --knn-only 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())['roofline']
print('$1: knn us',round(r['us_per_launch'],1),'min',round(r['us_min'],1),'local us',round(r['local_regime']['us_per_launch'],1))"; }
for nq in 16384 32768 65536 131072; do MB_BENCH_NQ=$nq run "nq=$nq"; done
cp mimosa_b200/lib/libmimosa_b200.so /tmp/lib_keep.so
cp mimosa_b200/lib/libmimosa_b200_skipnb.so mimosa_b200/lib/libmimosa_b200.so
run "skip-nb"
cp /tmp/lib_keep.so mimosa_b200/lib/libmimosa_b200.so
