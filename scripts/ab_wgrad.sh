#!/bin/bash
# weight gradient on the FP32 pipe (wg=0) vs on the tensor pipe (wg=1), tensor-pipe engine
run() { PDL_ENGINE=mma PDL_DGRAD_MMA=$1 PDL_WGRAD_MMA=$2 python scripts/probe_one.py $3 $4 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"mma dg=$1 wg=$2\", /"; }
for w in "heat Tanh" "kdv Tanh" "heat Rational"; do
  set -- $w
  run 1 0 $1 $2; run 1 1 $1 $2; run 0 1 $1 $2
done
for w in "ks Tanh" "rd2d Tanh" "ks Rational"; do
  set -- $w
  run 0 0 $1 $2; run 0 1 $1 $2
done
