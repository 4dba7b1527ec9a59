#!/bin/bash
# ncu capture of k_knn with a variant library.  $1 = variant suffix, $2 = tag
cp mimosa_b200/lib/libmimosa_b200.so /tmp/lib_keep.so
cp mimosa_b200/lib/libmimosa_b200_$1.so mimosa_b200/lib/libmimosa_b200.so
bash tools/gpu_prof_knn.sh $2
cp /tmp/lib_keep.so mimosa_b200/lib/libmimosa_b200.so
