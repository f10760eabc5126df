#!/bin/sh
# developer-only: host build of the per-sample functions (see hostsim.cpp)
D=$(dirname "$0")
g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -I"$D/../../kgeo_b200/csrc" "$D/hostsim.cpp" -o "$D/_hostsim.so"
