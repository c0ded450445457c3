"""CPU (no GPU): parity errors of the EMULATED kernel (tests/emu.py, test infrastructure) vs the reference's float64 golden vectors
with the current PDL_* knobs -- answers precision questions about a kernel variant before spending GPU time.
usage: PDL_CROSS_BF16=1 python scripts/emu_errors.py [case ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pde_learn_b200 import _lib as L
from tests import emu
from tests.helpers import GOLDEN_CASES, GoldenCase, rel_max
from tests.test_emulation_parity import desc_for

print("knobs:", {k: v for k, v in os.environ.items() if k.startswith("PDL_")})
worst = 0.0
for name in (sys.argv[1:] or GOLDEN_CASES):
    g = GoldenCase(name)
    desc, keep = desc_for(g)
    out = emu.run_emu(desc, g.points.numpy(), [p.numpy() for p in g.params], g.xi.numpy(), np.array(g.mask, dtype=np.uint8), n_ctas=3)
    ref = np.concatenate([x.numpy().ravel() for x in g.grads])
    src = L.plan_source(desc)
    jets = [tuple(int(v) for v in part.split("=(")[1].rstrip(")").split(",")) for part in src.splitlines()[1].replace("// jets:", "").split()]
    fe = max(rel_max(out["features"][jets.index(tuple(enc) + (0,) * (4 - len(enc)))], g.z["features"][j]) for j, enc in enumerate(g.encodings))
    want = float(g.z["coll_loss"])
    e = [abs(out["loss"] - want) / want, rel_max(out["residual"], g.z["residual"]), rel_max(out["grad_params"], ref),
         rel_max(out["grad_xi"], g.z["grad_xi"]), fe]
    worst = max(worst, max(e))
    print("%-22s loss %.1e resid %.1e dtheta %.1e dxi %.1e feats %.1e" % (name, *e), flush=True)
print("worst %.2e" % worst)
