#!/bin/bash
# Quick k-NN iteration: parity tests, then the k-NN-only timing for each MB_KNN_PREF value given.  $@ = values
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in "$@"; do
  MB_KNN_PREF=$v timeout 600 python -u bench.py --knn-only 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())['roofline']
print('pref $v: knn us',round(r['us_per_launch'],1),'min',round(r['us_min'],1),'frac',round(r['frac'],3),'local us',round(r['local_regime']['us_per_launch'],1))"
done
