#!/bin/sh
# builds tools/analyze_hash.bin from the host analysis sources
cd "$(dirname "$0")/.." && g++ -O2 -std=c++17 -pthread -o tools/analyze_hash.bin tools/analyze_hash.cpp madipm_jl_b200/csrc/ls_symbolic.cpp madipm_jl_b200/csrc/host_symbolic.cpp
