#!/bin/bash
# Fifth sweep: is a weight-gradient step of all 5 n8 tiles now the better choice everywhere (the L2-reduction flush freed registers)?
cd "$(dirname "$0")/.." || exit 1
export PDL_PROBE_FLUSH=1
run() { local label=$1 w=$2 a=$3; shift 3
  r=$(env "$@" timeout 150 python scripts/probe_one.py $w $a 2>/dev/null | tail -1)
  echo "$w $a [$label] $(echo "$r" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms (min %.4f)" % (d["ms"], d["min_ms"]))' 2>/dev/null || echo failed)"
}
for rep in 1 2; do
run "NTB=5" heat Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "NTB=5 TANH_CACHE=1" heat Tanh "PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5 -DPDL_TANH_CACHE=1"
run "default" ks Tanh X=1
run "NTB=5" ks Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" rd2d Tanh X=1
run "NTB=5" rd2d Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" heat Rational X=1
run "NTB=5" heat Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" ks Rational X=1
run "NTB=5" ks Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" kdv Rational X=1
run "NTB=5" kdv Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
done
