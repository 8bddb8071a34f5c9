#!/bin/sh
# Amalgamation sweep (MIPM_RELAX = always,k1,z1,k2,z2,z3) on the K2 system of C3 and on C2: factor / solve times.
for r in "16,64,0.5,128,0.3,0.1" "8,32,0.5,64,0.3,0.1" "32,96,0.6,192,0.4,0.15" "24,64,0.6,128,0.4,0.2" "16,64,0.3,128,0.2,0.05" "4,16,0.3,32,0.2,0.05"; do
  echo "== MIPM_RELAX=$r"
  MIPM_RELAX=$r python tools/profile_factor_k2.py 2>&1 | grep -E "nnz_l|^factor ms|^solve ms" | cut -c1-200
  MIPM_RELAX=$r python tools/profile_factor.py 2>&1 | grep -E "factorize \(events|solve \(events"
done
