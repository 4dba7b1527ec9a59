#!/bin/bash
# ncu launch list of the ICP kernels (2 scans)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"k_linearize|k_finalize|k_loc_comp|k_group_sort|k_copy_src|k_pack_src" -c 120 --csv --log-file gpurun_out/launches_icp.csv python -u bench.py --profile-icp --steps 2 > gpurun_out/ncu_icp.log 2>&1; echo ncu rc=$?
python tools/launch_summary.py gpurun_out/launches_icp.csv 50 2>&1 | tail -14
