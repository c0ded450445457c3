#!/bin/bash
# Third sweep: does the J = 5 optimum (5 n8 tiles per weight-gradient step) hold for Burgers (K = 17) and for Rational nets?
cd "$(dirname "$0")/.." || exit 1
export PDL_PROBE_FLUSH=1
run() { local label=$1 w=$2 a=$3; shift 3
  r=$(env "$@" timeout 150 python scripts/probe_one.py $w $a 2>/dev/null | tail -1)
  echo "$w $a [$label] $(echo "$r" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms (min %.4f)" % (d["ms"], d["min_ms"]))' 2>/dev/null || echo failed)"
}
for rep in 1 2; do
run "default" burgers Tanh X=1
run "NTB=5" burgers Tanh PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" kdv Rational X=1
run "NTB=5" kdv Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
run "default" burgers Rational X=1
run "NTB=5" burgers Rational PDL_NVCC_FLAGS=-DPDL_WGRAD_NTB=5
done
