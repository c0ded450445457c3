#!/bin/bash
# working tree kernels vs the copy of the previous kernels in experiments/_v1_csrc (same box, alternating)
for w in "heat Tanh" "kdv Tanh" "ks Tanh" "rd2d Tanh" "heat Rational"; do
  set -- $w
  for i in 1 2; do
    python scripts/probe_one.py $1 $2 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"new\", /" | cut -c1-150
    PDL_CSRC_DIR=$PWD/experiments/_v1_csrc python scripts/probe_one.py $1 $2 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"old\", /" | cut -c1-150
  done
done
