#!/bin/bash
# kernel-only sweep: "workload PPL NW UK UQ" (run on a B200)
for cfg in "heat 1 8 2 1" "heat 1 8 2 2" "heat 1 10 2 1" "heat 1 12 2 1" "heat 1 12 1 1" "heat 2 8 2 1" "heat 2 8 1 2" "ks 1 8 2 1" "ks 1 8 1 1" "ks 1 10 1 1"; do
  set -- $cfg
  PDL_PPL=$2 PDL_NW=$3 PDL_UK=$4 PDL_UQ=$5 python scripts/probe_one.py $1 Tanh 2>&1 | tail -1 | sed "s/^{/{\"cfg\": \"$cfg\", /"
done
