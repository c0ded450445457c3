#!/bin/bash
run() { PDL_NVCC_FLAGS="$5" PDL_ENGINE=mma PDL_DGRAD_MMA=$1 PDL_WGRAD_MMA=$2 python scripts/probe_one.py $3 $4 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"mma dg=$1 wg=$2 $5\", /"; }
run 2 1 ks Tanh ""; run 2 1 ks Tanh "-DPDL_DGRAD_SWAP"
run 2 1 rd2d Tanh ""; run 2 1 rd2d Tanh "-DPDL_DGRAD_SWAP"
run 2 0 ks Rational ""; run 2 0 ks Rational "-DPDL_DGRAD_SWAP"
