#!/bin/bash
cp mimosa_b200/lib/libmimosa_b200.so /tmp/lib_keep.so
cp mimosa_b200/lib/libmimosa_b200_lint.so mimosa_b200/lib/libmimosa_b200.so
timeout 600 python -u bench.py --value-only --steps 1 --no-graph 2>/dev/null | grep "lin timing" | sed -n '61,72p'
cp /tmp/lib_keep.so mimosa_b200/lib/libmimosa_b200.so
