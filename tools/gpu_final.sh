#!/bin/bash
# Round-end GPU captures (one GPU): launch list of two ICP steps, full ncu capture of k_icp_loop (whole 20-iteration
# loop launch) and of k_knn (spread queries), copied into gpurun_out/ for tools/profile_summaries.sh.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_icp_2steps.csv python -u bench.py --profile-icp --steps 2 > gpurun_out/ncu_icp.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_icp_loop -s 1 -c 1 -o gpurun_out/prof_r2_icp_loop python -u bench.py --profile-icp --steps 2 > gpurun_out/ncu_icp2.log 2>&1; echo "icp_loop full rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_knn -s 1 -c 1 -o gpurun_out/prof_r2_knn python -u bench.py --profile-knn --steps 3 > gpurun_out/ncu_knn.log 2>&1; echo "knn full rc=$?"
