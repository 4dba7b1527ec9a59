#!/bin/bash
cp mimosa_b200/lib/libmimosa_b200.so /tmp/lib_keep.so
cp mimosa_b200/lib/libmimosa_b200_timing.so mimosa_b200/lib/libmimosa_b200.so
for nq in 131072; do echo "nq=$nq"; MB_BENCH_NQ=$nq timeout 600 python -u bench.py --knn-only 2>/dev/null | grep "knn timing" | sed -n '3,9p'; done
cp /tmp/lib_keep.so mimosa_b200/lib/libmimosa_b200.so
