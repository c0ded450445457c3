#!/bin/bash
# does the 168-register code alone explain the slowdown at > 8 warps?  ("workload NW nvcc-flags")
PDL_NW=8 PDL_NVCC_FLAGS="-maxrregcount=168" python scripts/probe_one.py heat Tanh 2>&1 | tail -1 | sed 's/^{/{"cfg": "heat NW=8 maxreg168", /'
PDL_NW=8 PDL_NVCC_FLAGS="-maxrregcount=168" python scripts/probe_one.py ks Tanh 2>&1 | tail -1 | sed 's/^{/{"cfg": "ks NW=8 maxreg168", /'
PDL_NW=12 python scripts/probe_one.py heat Tanh 262144 2>&1 | tail -1 | sed 's/^{/{"cfg": "heat NW=12 B=2^18", /'
PDL_NW=8 python scripts/probe_one.py heat Tanh 262144 2>&1 | tail -1 | sed 's/^{/{"cfg": "heat NW=8 B=2^18", /'
