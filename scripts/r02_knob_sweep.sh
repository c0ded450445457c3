#!/bin/bash
# One-at-a-time variations of the adjoint knobs around the plan compiler's defaults, with the fp16 forward in place (the defaults were
# tuned in round 1 with a 3xTF32 forward).  Kernel-only ms at 2^20 rows, L2 flushed between launches.
cd "$(dirname "$0")/.." || exit 1
export PDL_PROBE_FLUSH=1
run() { # run <label> <workload> <act> [env assignments...]
  local label=$1 w=$2 a=$3; shift 3
  r=$(env "$@" timeout 150 python scripts/probe_one.py $w $a 2>/dev/null | tail -1)
  echo "$w $a [$label] $(echo "$r" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms (min %.4f)" % (d["ms"], d["min_ms"]))')"
}
for rep in 1 2; do
run "default" heat Tanh X=1
run "ADJ_FUSED=1" heat Tanh PDL_ADJ_FUSED=1
run "NTB=2" heat Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
run "NTB=5" heat Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" heat Rational X=1
run "ADJ_FUSED=0" heat Rational PDL_ADJ_FUSED=0
run "NTB=2" heat Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
run "NTB=5" heat Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" ks Tanh X=1
run "NTB=2" ks Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=2
run "NTB=5" ks Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "ADJ_FUSED=0" ks Tanh PDL_ADJ_FUSED=0
done
