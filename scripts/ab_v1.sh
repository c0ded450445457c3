#!/bin/bash
# A/B in one box: committed v1 kernel template (PDL_CSRC_DIR) vs the working-tree template
for w in "kdv Tanh" "ks Tanh" "heat Rational" "heat Tanh" "ks Rational" "rd2d Tanh"; do
  set -- $w
  PDL_CSRC_DIR=$PWD/experiments/_v1_csrc python scripts/probe_one.py $1 $2 2>&1 | tail -1 | sed 's/^{/{"variant": "v1", /'
  python scripts/probe_one.py $1 $2 2>&1 | tail -1 | sed 's/^{/{"variant": "new", /'
done
