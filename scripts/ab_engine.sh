#!/bin/bash
# FP32-pipe engine vs the default tensor-pipe engine (plan compiler's own choice of dgrad/wgrad modes), kernel ms at 2^20 rows
run() { PDL_ENGINE=$1 python scripts/probe_one.py $2 $3 1048576 $4 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$1 $4\", /"; }
for w in "heat Tanh" "burgers Tanh" "kdv Tanh" "ks Tanh" "rd2d Tanh" "heat Rational" "burgers Rational" "ks Rational" "rd2d Rational"; do
  set -- $w
  run ffma $1 $2; run mma $1 $2
done
run ffma heat Tanh fwd; run mma heat Tanh fwd; run ffma ks Tanh fwd; run mma ks Tanh fwd
