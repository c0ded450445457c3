#!/bin/bash
# FP32-pipe kernel vs tensor-pipe engine (component pairs as m16 row halves, 8-point tiles)
run() { PDL_ENGINE=$1 PDL_DGRAD_MMA=$2 python scripts/probe_one.py $3 $4 1048576 $5 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$1 dg=$2 $5\", /"; }
for w in "heat Tanh" "kdv Tanh" "heat Rational"; do
  set -- $w
  run ffma 0 $1 $2; run mma 1 $1 $2; run mma 0 $1 $2
done
for w in "ks Tanh" "rd2d Tanh" "ks Rational"; do
  set -- $w
  run ffma 0 $1 $2; run mma 0 $1 $2
done
run ffma 0 heat Tanh fwd; run mma 0 heat Tanh fwd; run ffma 0 ks Tanh fwd; run mma 0 ks Tanh fwd
