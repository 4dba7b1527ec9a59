#!/bin/bash
# memcheck (+ racecheck with $1 = race) of the C++ mirror smoke program (tiny map, linearize calls, knn)
cd tests/cpp && g++ -std=c++17 -O1 host_mirror_smoke.cpp -I ../../mimosa_b200/host -L ../../mimosa_b200/lib -lmimosa_b200 -Wl,-rpath,$PWD/../../mimosa_b200/lib -o host_mirror_smoke || exit 1
timeout 150 compute-sanitizer --tool memcheck ./host_mirror_smoke 2>&1 | grep -v "Host Frame\|Saved host" | tail -12
if [ "$1" = race ]; then timeout 150 compute-sanitizer --tool racecheck ./host_mirror_smoke 2>&1 | grep -v "Host Frame\|Saved host" | tail -12; fi
