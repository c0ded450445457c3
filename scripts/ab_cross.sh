#!/bin/bash
# 3xTF32 (mask 0) vs bf16 cross-term MMA (PDL_CROSS_BF16 bit mask: 1 forward, 2 data gradient, 4 weight gradient), kernel ms at 2^20 rows
run() { PDL_CROSS_BF16=$1 PDL_NVCC_FLAGS="$2" python scripts/probe_one.py $3 $4 1048576 $5 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"mask $1 $2\", /"; }
for w in "heat Tanh" "kdv Tanh" "ks Tanh" "rd2d Tanh" "heat Rational" "ks Rational"; do
  set -- $w
  for m in 0 6 7; do run $m "" $1 $2; done
done
for w in "heat Tanh" "ks Tanh"; do
  set -- $w
  for m in 6 7; do run $m "-DPDL_XACC=1" $1 $2; done
done
run 0 "" heat Tanh fwd; run 7 "" heat Tanh fwd; run 7 "-DPDL_XACC=1" heat Tanh fwd
for m in 6 7; do PDL_CROSS_BF16=$m python scripts/gpu_errors.py 2>&1 | tail -13; done
