#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u tools/iter_times.py 2>&1 | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_icp.csv python -u bench.py --profile-icp --steps 2 > gpurun_out/ncu_icp.log 2>&1; echo ncu rc=$?
python tools/launch_summary.py gpurun_out/launches_icp.csv 60 2>&1 | tail -30
