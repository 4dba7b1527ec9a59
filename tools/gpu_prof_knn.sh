#!/bin/bash
# One full ncu capture of k_knn on the spread query set.  $1 = tag
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn -s 1 -c 1 -o gpurun_out/prof_knn_$TAG -f python -u bench.py --profile-knn --steps 3 > gpurun_out/ncu_knn.log 2>&1; echo ncu rc=$?
