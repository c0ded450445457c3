"""Shared test helpers: golden-case loading and conversion to oracle inputs.  Test infrastructure only."""
import json
import os

import numpy as np
import torch

from oracle import pde_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_CASES = ["burgers_tanh", "burgers_tanh_masked", "burgers_rational20", "heat_tanh", "heat_rational",
                "kdv_tanh", "ks_tanh", "ks_rational", "rd2d_tanh", "wave_mixed_tanh", "heat3d_tanh",
                # networks / Xi from the END of reference training runs on real data (tests/golden/make_golden_trained.py)
                "burgers_tanh20_trained", "ks_tanh_trained", "ks_rational_trained"]


def _spec(blob):
    s = json.loads(str(blob))
    fix = lambda t: [(tuple(e), int(p)) for e, p in t]
    return fix(s["lhs"]), [fix(t) for t in s["rhs"]]


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.z = z
        self.d = int(z["d"])
        self.widths = [int(w) for w in z["widths"]]
        self.activation = str(z["activation"])
        self.lhs, self.rhs = _spec(z["spec"])
        self.encodings = O.distinct_encodings(self.lhs, self.rhs)
        self.n_params = int(z["n_params"])
        self.params = [torch.from_numpy(z["param_%02d" % i].copy()) for i in range(self.n_params)]
        self.grads = [torch.from_numpy(z["grad_%02d" % i].copy()) for i in range(self.n_params)]
        self.dgrads = [torch.from_numpy(z["dgrad_%02d" % i].copy()) for i in range(self.n_params)]
        self.xi = torch.from_numpy(z["xi"].copy())
        self.mask = [bool(m) for m in z["mask"]]
        self.points = torch.from_numpy(z["points"].copy())
        self.x_data = torch.from_numpy(z["x_data"].copy())
        self.y_data = torch.from_numpy(z["y_data"].copy())
        self.p = float(z["p"])
        # trained cases record how far the reference's OWN float32 path is from its float64 path (make_golden.py: ref_fp32_err)
        self.ref_fp32_err = dict(zip(("loss", "resid", "dtheta", "dxi", "feats"), [float(v) for v in z["ref_fp32_err"]])) \
            if "ref_fp32_err" in z.files else None

    def tol(self, quantity, base=1e-5):
        """Parity gate for `quantity` in (loss, resid, dtheta, dxi, feats): the north star's 1e-5 -- except where the case is so
        ill-conditioned that the reference's own float32 evaluation misses its float64 one by more than a quarter of that; there the
        gate is 4x the reference's float32 error (a float32 kernel cannot be asked to beat float32 arithmetic by much)."""
        if self.ref_fp32_err is None:
            return base
        return max(base, 4.0 * self.ref_fp32_err[quantity])

    def net(self, dtype=torch.float64, requires_grad=True):
        """Oracle NetParams from the stored float32 parameters (reference parameter order)."""
        nl = len(self.widths) - 1
        Ws = [self.params[2 * i] for i in range(nl)]
        bs = [self.params[2 * i + 1] for i in range(nl)]
        ra = rb = []
        if self.activation == "rational":
            ra = [self.params[2 * nl + 2 * i] for i in range(nl - 1)]
            rb = [self.params[2 * nl + 2 * i + 1] for i in range(nl - 1)]
        return O.NetParams(Ws, bs, self.activation, list(ra), list(rb)).to(dtype, requires_grad)


def rel_max(a, b):
    """max|a-b| / max|b| (the SURVEY 8d 'relative max-norm' gate)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
