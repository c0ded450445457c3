#!/bin/bash
for w in "heat Tanh" "heat Rational"; do
  set -- $w
  PDL_ENGINE=ffma python scripts/probe_one.py $1 $2 2>&1 | tail -1 | sed 's/^{/{"variant": "ffma", /'
  PDL_ENGINE=mma PDL_DGRAD_MMA=1 python scripts/probe_one.py $1 $2 2>&1 | tail -1 | sed 's/^{/{"variant": "mma fwd+dgrad", /'
  PDL_ENGINE=mma PDL_DGRAD_MMA=0 python scripts/probe_one.py $1 $2 2>&1 | tail -1 | sed 's/^{/{"variant": "mma fwd only", /'
done
