#!/bin/bash
# Fourth sweep: heat tanh again, now that the L2-reduction flush freed 19 registers (236, no spills).
cd "$(dirname "$0")/.." || exit 1
export PDL_PROBE_FLUSH=1
run() { local label=$1 w=$2 a=$3; shift 3
  r=$(env "$@" timeout 150 python scripts/probe_one.py $w $a 2>/dev/null | tail -1)
  echo "$w $a [$label] $(echo "$r" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms (min %.4f)" % (d["ms"], d["min_ms"]))' 2>/dev/null || echo failed)"
}
for rep in 1 2 3; do
run "default (NTB=2)" heat Tanh X=1
run "NTB=3" heat Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=3
run "NTB=5" heat Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "TANH_CACHE=1" heat Tanh PDL_NVCC_FLAGS=-DPDL_TANH_CACHE=1
run "ADJ_FUSED=1" heat Tanh PDL_ADJ_FUSED=1
run "ADJ_FUSED=1 NTB=3" heat Tanh PDL_ADJ_FUSED=1 PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=3
run "ROLL_MT=1 NTB=3" heat Tanh "PDL_NVCC_FLAGS=-DPDL_WGRAD_ROLL_MT=1 -DPDL_WGRAD_NTB=3"
done
