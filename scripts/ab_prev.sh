#!/bin/bash
# same-box A/B of the working-tree kernel against a previous copy of the kernel template under experiments/_prev_csrc (PDL_CSRC_DIR)
run() { PDL_CSRC_DIR=$1 python scripts/probe_one.py $2 $3 1048576 $4 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"${1:+prev}\", /"; }
for w in "heat Tanh" "burgers Tanh" "kdv Tanh" "ks Tanh" "rd2d Tanh" "heat Rational" "burgers Rational" "ks Rational" "rd2d Rational"; do
  set -- $w
  run /root/repo/experiments/_prev_csrc $1 $2; run "" $1 $2
done
