#!/bin/bash
run() { timeout 600 python -u bench.py --value-only --steps 10 2>/dev/null | tail -1; }
echo "lb5:"; run
cp mimosa_b200/lib/libmimosa_b200.so /tmp/lib_keep.so
for v in "$@"; do cp mimosa_b200/lib/libmimosa_b200_$v.so mimosa_b200/lib/libmimosa_b200.so; echo "$v:"; run; done
cp /tmp/lib_keep.so mimosa_b200/lib/libmimosa_b200.so
