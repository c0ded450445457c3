"""One kernel-only probe: python scripts/probe_one.py WORKLOAD ACT [ROWS] (honours PDL_PPL / PDL_NW)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.perf_probe import probe

if __name__ == "__main__":
    w, act = sys.argv[1], sys.argv[2]
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    bwd = not (len(sys.argv) > 4 and sys.argv[4] == 'fwd')
    r = probe(w, act, B, bwd=bwd)
    r["PPL"] = os.environ.get("PDL_PPL", "default"); r["NW"] = os.environ.get("PDL_NW", "default")
    print(json.dumps(r), flush=True)
