"""Prints the measured parity errors (relative max-norm vs the reference's float64 golden vectors) of the current engine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pde_learn_b200 as P
from tests.helpers import GOLDEN_CASES, GoldenCase, rel_max
from tests.test_gpu_parity import _network_from_golden, _terms

print("engine:", os.environ.get("PDL_ENGINE", "default"))
worst = 0.0
for name in GOLDEN_CASES:
    g = GoldenCase(name); U = _network_from_golden(g, P); derivs, lhs, rhs = _terms(P, g.lhs, g.rhs)
    xi = g.xi.to("cuda").requires_grad_(True)
    loss, r = P.Coll_Loss(U, xi, torch.tensor(g.mask), g.points.to("cuda"), derivs, lhs, rhs); loss.backward()
    got = np.concatenate([p.grad.cpu().numpy().ravel() for p in U.parameters()])
    ref = np.concatenate([x.numpy().ravel() for x in g.grads])
    f = P.Evaluate_Derivatives(U, derivs, g.points.to("cuda")).cpu().numpy()
    e = [abs(float(loss.detach()) - float(g.z["coll_loss"])) / float(g.z["coll_loss"]), rel_max(r.cpu().numpy(), g.z["residual"]),
         rel_max(got, ref), rel_max(xi.grad.cpu().numpy(), g.z["grad_xi"]), max(rel_max(f[j], g.z["features"][j]) for j in range(len(derivs)))]
    worst = max(worst, max(e))
    print("%-22s loss %.1e resid %.1e dtheta %.1e dxi %.1e feats %.1e" % (name, *e))
print("worst %.2e" % worst)
