#!/bin/bash
run() { PDL_NW=$7 PDL_NVCC_FLAGS="$5" PDL_ENGINE=mma PDL_DGRAD_MMA=$1 PDL_WGRAD_MMA=$2 python scripts/probe_one.py $3 $4 1048576 $6 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"mma dg=$1 wg=$2 $5 $6 nw=$7\", /"; }
run 0 0 ks Rational "" "" ""
run 0 1 ks Rational "" "" ""
run 0 1 heat Rational "" "" ""
run 1 0 burgers Rational "" "" ""
run 1 1 burgers Rational "" "" ""
run 0 0 rd2d Tanh "" "" ""
run 0 1 rd2d Tanh "" "" ""
run 0 0 rd2d Rational "" "" ""
run 0 1 rd2d Rational "" "" ""
