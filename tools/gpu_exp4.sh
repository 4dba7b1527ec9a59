#!/bin/bash
run() { timeout 600 python -u bench.py --knn-only 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())['roofline']
print('$1: knn us',round(r['us_per_launch'],1),'min',round(r['us_min'],1),'local us',round(r['local_regime']['us_per_launch'],1))"; }
run "cold code"
MB_BENCH_WARM_CODE=32 run "warm code (32-query launch after the flush)"
MB_BENCH_WARM_CODE=4736 run "warm code (4736 = one block per SM)"
