#!/bin/bash
# development: per-iteration times of the kernel variants
for kern in v1 v3; do for sort in 1 0; do
  export MB_LIN_KERNEL=$kern MB_LIN_SORT=$sort
  echo "== kernel=$kern sort=$sort"; timeout 300 python tools/iter_times.py 2>&1 | tail -8
done; done
MB_LIN_KERNEL=v1 MB_LIN_SORT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
