#!/bin/bash
# ncu launch list of the ICP kernels (2 scans) + full capture of the first 7 k_linearize launches of a scan
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_linearize|k_finalize|k_loc_comp|k_sort_keys|k_apply_perm|RadixSort" -c 200 --csv --log-file gpurun_out/launches_icp.csv python -u bench.py --profile-icp --steps 2 > gpurun_out/ncu_icp.log 2>&1; echo ncu rc=$?
python tools/launch_summary.py gpurun_out/launches_icp.csv 40 2>&1 | tail -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_linearize -c 7 -f -o gpurun_out/prof_lin python -u bench.py --profile-icp --steps 1 > gpurun_out/ncu_lin.log 2>&1; echo ncu2 rc=$?
ls -la gpurun_out/*.ncu-rep
