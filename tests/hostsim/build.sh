#!/bin/sh
# TEST-ONLY: host build of the per-lane solver code for CPU-side debugging (see hostsim.cpp header).
cd "$(dirname "$0")" && g++ -O2 -std=c++17 -fPIC -shared -Wno-unknown-pragmas -o libhostsim.so hostsim.cpp
