#!/bin/bash
# Second sweep (see r02_knob_sweep.sh): the remaining knobs and workloads.
cd "$(dirname "$0")/.." || exit 1
export PDL_PROBE_FLUSH=1
run() { local label=$1 w=$2 a=$3; shift 3
  r=$(env "$@" timeout 150 python scripts/probe_one.py $w $a 2>/dev/null | tail -1)
  echo "$w $a [$label] $(echo "$r" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms (min %.4f)" % (d["ms"], d["min_ms"]))' 2>/dev/null || echo failed)"
}
for rep in 1 2; do
run "default" heat Tanh X=1
run "ROLL_MT=1" heat Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_ROLL_MT=1
run "TANH_CACHE=1" heat Tanh PDL_NVCC_FLAGS=-DPDL_TANH_CACHE=1
run "NTB=1" heat Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=1
run "NW=7" heat Tanh PDL_NW=7
run "default" kdv Tanh X=1
run "NTB=2" kdv Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
run "NTB=5" kdv Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "DGRAD_MMA=2" kdv Tanh PDL_DGRAD_MMA=2
run "default" rd2d Tanh X=1
run "NTB=2" rd2d Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
run "NTB=5" rd2d Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" ks Rational X=1
run "NTB=2" ks Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
run "ROLL_MT=0" heat Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_ROLL_MT=0
run "default" heat Rational X=1
done
