#!/bin/bash
cd tests/cpp && g++ -std=c++17 -O1 host_mirror_smoke.cpp -I ../../mimosa_b200/host -L ../../mimosa_b200/lib -lmimosa_b200 -Wl,-rpath,$PWD/../../mimosa_b200/lib -o host_mirror_smoke || exit 1
timeout 600 compute-sanitizer --tool racecheck --racecheck-report hazard ./host_mirror_smoke 2>&1 | grep -v "Host Frame\|Saved host" | head -60
