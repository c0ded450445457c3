#!/bin/bash
# A/B: weight-gradient partial sums flushed with one vector reduction at the L2 (PDL_WGRAD_RED=1) instead of load + add + store.
cd "$(dirname "$0")/.." || exit 1
export PDL_PROBE_FLUSH=1
for f in "" "-DPDL_WGRAD_RED=1" ""; do
  echo "== flags: $f"
  PDL_NVCC_FLAGS="$f" timeout 400 python scripts/perf_probe.py 2>&1 | grep "^{" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if 'ms' in d and d['B'] > 100000: print('  %-8s %-8s bwd=%d  %.4f ms (min %.4f)' % (d['workload'], d['act'], d['bwd'], d['ms'], d['min_ms']))
"
done
echo "== parity with -DPDL_WGRAD_RED=1"
PDL_NVCC_FLAGS="-DPDL_WGRAD_RED=1" timeout 300 python scripts/gpu_errors.py 2>&1 | grep -v Warn | grep "loss\|worst"
