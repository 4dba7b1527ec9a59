#!/bin/bash
# Development GPU round: parity tests, then the device-resident loop timing with and without bulk staging.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for pool in default 0; do
  if [ $pool = default ]; then unset MB_LIN_POOL; else export MB_LIN_POOL=$pool; fi
  timeout 600 python -u bench.py --value-only --steps 10 --warmup 3 > gpurun_out/bench_value_pool_$pool.log 2>&1; echo "bench pool=$pool rc=$?"
  tail -2 gpurun_out/bench_value_pool_$pool.log | cut -c1-600
done
