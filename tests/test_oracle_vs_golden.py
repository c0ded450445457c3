"""
Pins oracle/pde_oracle.py (CPU, no GPU needed):
  (a) against golden vectors produced by the unmodified reference (tests/golden/make_golden.py);
  (b) against the known answers of the reference's own unit tests.
"""
import json
import math
import os
import random

import numpy as np
import pytest
import torch

from oracle import pde_oracle as O
from tests.helpers import GOLDEN_CASES, GOLDEN_DIR, GoldenCase, rel_max

TOL = 1e-12          # oracle and reference are both float64 autograd: agreement to rounding


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_coll_loss_matches_reference(name):
    g = GoldenCase(name)
    net = g.net()
    xi = g.xi.double().requires_grad_(True)
    P = g.points.double()
    loss, resid, feats = O.coll_loss(lambda X: O.mlp_forward(net, X), xi, g.mask, P, g.lhs, g.rhs)
    assert abs(float(loss.detach()) - float(g.z["coll_loss"])) <= TOL * max(1.0, abs(float(g.z["coll_loss"])))
    assert rel_max(resid.detach().numpy(), g.z["residual"]) < 1e-11
    for j, enc in enumerate(g.encodings):
        assert rel_max(feats[enc].detach().numpy(), g.z["features"][j]) < 1e-11, enc
    grads = torch.autograd.grad(loss, net.tensors() + [xi])
    for got, want in zip(grads[:-1], g.grads):
        assert rel_max(got.numpy(), want.numpy()) < 1e-10
    assert rel_max(grads[-1].numpy(), g.z["grad_xi"]) < 1e-10
    for k, m in enumerate(g.mask):
        if m:
            assert float(grads[-1][k]) == 0.0


@pytest.mark.parametrize("name", ["burgers_tanh", "heat_rational", "rd2d_tanh"])
def test_data_lp_l2_match_reference(name):
    g = GoldenCase(name)
    net = g.net()
    d = O.data_loss(lambda X: O.mlp_forward(net, X), g.x_data.double(), g.y_data)
    assert abs(float(d.detach()) - float(g.z["data_loss"])) < 1e-12
    dg = torch.autograd.grad(d, net.tensors())
    for got, want in zip(dg, g.dgrads):
        assert rel_max(got.numpy(), want.numpy()) < 1e-10
    xi = g.xi.double().requires_grad_(True)
    lp = O.lp_loss(xi, g.mask, g.p)
    assert abs(float(lp) - float(g.z["lp_loss"])) < 1e-12
    assert rel_max(torch.autograd.grad(lp, [xi])[0].numpy(), g.z["lp_grad"]) < 1e-12
    # the reference accumulates L2 in a float32 [1] buffer (Loss.py:344) => its value is f32-rounded
    assert abs(float(O.l2_squared_loss(net)) - float(g.z["l2_loss"])) < 2e-6 * float(g.z["l2_loss"])


def test_config_library_matches_reference_reader():
    """config_library('burgers') == what the reference's Read_Library makes of the shipped Library.txt."""
    with open(os.path.join(GOLDEN_DIR, "library_txt_parsed.json")) as f:
        blob = json.load(f)
    d, lhs, rhs, _ = O.config_library("burgers")
    conv = lambda t: [[list(e), p] for e, p in t]
    assert conv(lhs) == blob["lhs"]
    assert [conv(t) for t in rhs] == blob["rhs"]
    assert [list(e) for e in O.distinct_encodings(lhs, rhs)] == blob["derivatives"]


def test_config_library_sizes():
    for name, J, K in [("burgers", 5, 17), ("heat", 4, 9), ("kdv", 5, 14), ("ks", 6, 30), ("rd2d", 6, 21)]:
        d, lhs, rhs, bounds = O.config_library(name)
        assert len(rhs) == K and len(O.distinct_encodings(lhs, rhs)) == J and len(bounds) == d


# ---- known answers of the reference's own tests ------------------------------------------------

def _poly2d(n):
    def f(X):
        t, x = X[:, 0], X[:, 1]
        return sum(t ** (n - i) * x ** i for i in range(n + 1))
    return f


def test_polynomial_derivatives_kat():
    """Test/Test_Evaluate_Derivatives.py:33-195: n=3 polynomial, analytic derivatives, abs tol 1e-5."""
    torch.manual_seed(0)
    X = (torch.rand(50, 2, dtype=torch.float64) * 2 - 1)
    t, x = X[:, 0], X[:, 1]
    f = O.derivative_features(_poly2d(3), X.clone(), [(0, 1), (1, 0), (0, 2), (2, 0), (1, 1)])
    want = {(0, 1): t ** 2 + 2 * t * x + 3 * x ** 2, (1, 0): 3 * t ** 2 + 2 * t * x + x ** 2,
            (0, 2): 2 * t + 6 * x, (2, 0): 6 * t + 2 * x, (1, 1): 2 * t + 2 * x}
    for enc, w in want.items():
        assert float((f[enc].detach() - w).abs().max()) < 1e-5


def test_coll_loss_kat_polynomial():
    """Test/Test_Loss.py:37-110: hand-assembled loss for U^4,(U_x)^3,(U_xx)^2,U_xxx,U(U_x)^2 on an n=4 polynomial."""
    torch.manual_seed(1)
    X = torch.rand(1000, 2, dtype=torch.float64); X[:, 1] = X[:, 1] * 2 - 1
    t, x = X[:, 0], X[:, 1]
    P = _poly2d(4)
    U, Ut, Ux, Uxx, Uxxx = (0, 0), (1, 0), (0, 1), (0, 2), (0, 3)
    rhs = [[(U, 4)], [(Ux, 3)], [(Uxx, 2)], [(Uxxx, 1)], [(U, 1), (Ux, 2)]]
    xi = torch.rand(5, dtype=torch.float64)
    loss, _, _ = O.coll_loss(P, xi, [False] * 5, X.clone(), [(Ut, 1)], rhs)
    u = P(X)
    ut = 4 * t ** 3 + 3 * t ** 2 * x + 2 * t * x ** 2 + x ** 3
    ux = t ** 3 + 2 * t ** 2 * x + 3 * t * x ** 2 + 4 * x ** 3
    uxx = 2 * t ** 2 + 6 * t * x + 12 * x ** 2
    uxxx = 6 * t + 24 * x
    pred = ((ut - (xi[0] * u ** 4 + xi[1] * ux ** 3 + xi[2] * uxx ** 2 + xi[3] * uxxx + xi[4] * u * ux ** 2)) ** 2).mean()
    assert abs(float(loss) - float(pred)) <= 1e-12 * abs(float(pred))


def test_lp_loss_kat():
    """Test/Test_Loss.py:180-223: xi=0 -> 0; xi==x -> N x^p; first M equal, rest 0 -> M x^p."""
    rnd = random.Random(3)
    N = rnd.randrange(5, 100); p = rnd.uniform(0.01, 0.1); x = rnd.uniform(0.01, 0.1)
    mask = [False] * N
    assert abs(float(O.lp_loss(torch.zeros(N, dtype=torch.float64), mask, p))) < 1e-5
    xi = torch.full((N,), x, dtype=torch.float64)
    assert abs(float(O.lp_loss(xi, mask, p)) - N * x ** p) < 1e-5
    M = rnd.randrange(1, N - 1); xi[M:] = 0
    assert abs(float(O.lp_loss(xi, mask, p)) - M * x ** p) < 1e-5
    with pytest.raises(AssertionError):
        O.lp_loss(xi, mask, 2.5)


def test_closure_matches_reference_training_step():
    """closure_losses reproduces the per-data-set losses the reference's Training() reported (float32 run)."""
    z = np.load(os.path.join(GOLDEN_DIR, "training_step_heat.npz"))
    s = json.loads(str(z["spec"]))
    fix = lambda t: [(tuple(e), int(p)) for e, p in t]
    lhs, rhs = fix(s["lhs"]), [fix(t) for t in s["rhs"]]
    nets = []
    for i in range(2):
        ps = [torch.from_numpy(z["p0_%d_%02d" % (i, j)].copy()) for j in range(int(z["n_params"]))]
        nl = len(ps) // 2
        nets.append(O.NetParams([ps[2 * k] for k in range(nl)], [ps[2 * k + 1] for k in range(nl)], "tanh")
                    .to(torch.float32, True))
    xi = torch.from_numpy(z["xi0"].copy()).requires_grad_(True)
    W = json.loads(str(z["weights"]))
    res = O.closure_losses(nets, xi, [bool(m) for m in z["mask"]],
                           [torch.from_numpy(z["coll_%d" % i].copy()) for i in range(2)],
                           [torch.from_numpy(z["inp_%d" % i].copy()) for i in range(2)],
                           [torch.from_numpy(z["tgt_%d" % i].copy()) for i in range(2)],
                           lhs, rhs, float(z["p"]), W)
    assert np.allclose(res.coll, z["train_Coll_Losses"], rtol=2e-5)
    assert np.allclose(res.data, z["train_Data_Losses"], rtol=2e-6)
    assert np.allclose(res.l2, z["train_L2_Losses"], rtol=2e-6)
    assert np.allclose(res.totals, z["train_Total_Losses"], rtol=2e-6)
    assert abs(res.lp - float(z["test_Lp_Loss"])) < 1e-6
    for i in range(2):
        assert rel_max(res.residuals[i].numpy(), z["resid_%d" % i]) < 5e-5
    # Adam's first step moves every parameter by lr*sign(grad) (up to eps): check direction on xi.
    moved = z["xi1"] - z["xi0"]
    g = res.grad_xi.numpy()
    nz = np.abs(g) > 1e-8
    assert np.all(np.sign(moved[nz]) == -np.sign(g[nz]))
    assert moved[3] == 0.0                         # masked component: zero gradient => Adam leaves it


def test_targeted_cutoff_select():
    torch.manual_seed(5)
    pts = torch.rand(300, 2); r = torch.randn(300); r[7] = 9.0; r[250] = -12.0
    cutoff, sel = O.targeted_cutoff_select(pts, r, 200)
    a = r.abs()
    assert abs(cutoff - float(a[:200].mean() + 3 * a[:200].std())) < 1e-6
    assert sel.shape[0] == int((a >= cutoff).sum()) and sel.shape[0] >= 2


def test_philox_oracle_known_answers():
    """The sampler's block function (oracle restatement) against Random123's published known-answer vectors."""
    for c, k, want in O.PHILOX_KAT:
        got = O.philox4x32_10(np.array([c], dtype=np.uint32), np.array([k], dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want


def test_sampler_oracle_properties():
    b = [(0.0, 10.0), (-8.0, 8.0)]
    a = O.sample_points(b, 4096, seed=7, first_index=0)
    s = O.sample_points(b, 1024, seed=7, first_index=1024)
    assert np.array_equal(a[1024:2048], s)                      # a shard is a slice of the global batch
    assert a.dtype == np.float32 and (a[:, 0] >= 0).all() and (a[:, 0] <= 10).all() and (a[:, 1] >= -8).all() and (a[:, 1] <= 8).all()
    assert abs(a[:, 0].mean() - 5.0) < 0.2 and abs(a[:, 1].mean()) < 0.3
