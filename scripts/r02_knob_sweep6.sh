#!/bin/bash
# Sixth sweep: the J = 6 and Rational plans once more after the L2-reduction flush (weight-gradient steps of 2 and 4 tiles).
cd "$(dirname "$0")/.." || exit 1
export PDL_PROBE_FLUSH=1
run() { local label=$1 w=$2 a=$3; shift 3
  r=$(env "$@" timeout 150 python scripts/probe_one.py $w $a 2>/dev/null | tail -1)
  echo "$w $a [$label] $(echo "$r" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms (min %.4f)" % (d["ms"], d["min_ms"]))' 2>/dev/null || echo failed)"
}
for rep in 1 2; do
run "default" ks Tanh X=1
run "NTB=2" ks Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
run "NTB=4" ks Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=4
run "default" rd2d Tanh X=1
run "NTB=2" rd2d Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
run "NTB=4" rd2d Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=4
run "default" heat Rational X=1
run "NTB=4" heat Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=4
run "NTB=2" heat Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
done
