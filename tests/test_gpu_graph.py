"""Captured_Epochs (the whole epoch as one CUDA graph) against the eager Run_Epochs loop."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(P, act="Tanh", S=1, seed=0):
    from pde_learn_b200 import configs
    d, bounds, derivs, lhs, rhs = configs.config("burgers")
    dev = torch.device("cuda", 0)
    torch.manual_seed(seed)
    U = [P.Network([d, 20, 20, 20, 1], act, Device=dev) for _ in range(S)]
    xi = (0.05 * torch.randn(len(rhs))).to(dev).requires_grad_(True)
    mask = torch.zeros(len(rhs), dtype=torch.bool)
    mask[3] = True
    g = torch.Generator().manual_seed(5)
    X = [torch.rand(300, d, generator=g).to(dev) * torch.tensor(bounds[:, 1] - bounds[:, 0], dtype=torch.float32, device=dev)
         + torch.tensor(bounds[:, 0], dtype=torch.float32, device=dev) for _ in range(S)]
    Y = [torch.sin(x[:, 1].double()) * torch.exp(-0.1 * x[:, 0].double()) for x in X]
    W = {"Data": 1.0, "Coll": 1.0, "Lp": 2e-4, "L2": 5e-6}
    return dev, U, xi, mask, X, Y, [bounds] * S, derivs, lhs, rhs, W


@pytest.mark.parametrize("act,S", [("Tanh", 1), ("Rational", 2)])
def test_captured_epochs_match_eager_loop(act, S):
    import pde_learn_b200 as P
    dev, U, xi, mask, X, Y, B, derivs, lhs, rhs, W = _setup(P, act, S)
    U2 = copy.deepcopy(U); xi2 = xi.detach().clone().requires_grad_(True)
    E, N = 12, 2000          # short on purpose: Adam amplifies last-bit differences (fused vs torch update) by ~1.1x per epoch
    opt = torch.optim.Adam([q for u in U for q in u.parameters()] + [xi], lr=2e-3)
    h1 = P.Run_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, opt, E, N, Seed=3)
    opt2 = torch.optim.Adam([q for u in U2 for q in u.parameters()] + [xi2], lr=2e-3)
    h2 = P.Run_Epochs(U2, xi2, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, opt2, E, N, Seed=3, Use_CUDA_Graph=True)
    # first epochs: same numbers (the fused Adam and torch's differ in the last bit, Adam then amplifies that by ~1.1x per epoch,
    # fastest in Lp = sum |xi|^p with p = 0.1, which is very steep near xi = 0)
    for key in ("Train Coll", "Train Data", "Train Total", "L2", "Lp"):
        np.testing.assert_allclose(h2[key][:2], h1[key][:2], rtol=2e-5, atol=1e-12, err_msg=key)
    for key in ("Train Coll", "Train Data", "Train Total", "L2"):
        np.testing.assert_allclose(h2[key], h1[key], rtol=1e-3, atol=1e-9, err_msg=key)
    np.testing.assert_allclose(h2["Lp"], h1["Lp"], rtol=3e-2, err_msg="Lp")
    assert np.abs(h2["Num Targeted"] - h1["Num Targeted"]).max() <= 2          # borderline rows may flip
    assert h1["Num Targeted"].max() > 0
    assert float((xi2 - xi).detach().abs().max()) <= 5e-4          # components crossing 0 feel the steep L^p weight |xi|^(p-2)
    assert float((xi2 - xi).detach().abs().median()) <= 1e-5
    assert float(xi2[3].detach()) == float(xi[3].detach())                                        # masked coefficient never moves
    for a, b in zip([q for u in U2 for q in u.parameters()], [q for u in U for q in u.parameters()]):
        assert float((a - b).detach().abs().max()) <= 2e-4
    # the optimiser keeps torch.optim.Adam's state layout (Code/main.py:519)
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys()
    for k in s1:
        assert float(s2[k]["step"]) == float(s1[k]["step"]) == E
        diff = (s2[k]["exp_avg"] - s1[k]["exp_avg"]).abs()
        scale = max(1e-30, float(s1[k]["exp_avg"].abs().max()))
        if k == max(s1.keys()):                       # Xi: the components near 0 see the L^p gradient ~ |xi|^(p-1), compare the bulk
            assert float(diff.median()) / scale <= 2e-2, (k, float(diff.median()) / scale)
        else:
            assert float(diff.max()) / scale <= 2e-2, (k, float(diff.max()) / scale)
        assert s2[k]["exp_avg_sq"].shape == s1[k]["exp_avg_sq"].shape


def test_captured_epochs_resume_and_restrictions():
    import pde_learn_b200 as P
    dev, U, xi, mask, X, Y, B, derivs, lhs, rhs, W = _setup(P)
    opt = torch.optim.Adam([q for u in U for q in u.parameters()] + [xi], lr=1e-3)
    cap = P.Captured_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, opt, 1000, Seed=1)
    x0 = xi.detach().clone()
    cap.run(3); cap.run(4)
    h = cap.history()
    assert h["Lp"].shape == (7,) and h["Train Coll"].shape == (7, 1)
    assert float((xi.detach() - x0).abs().max()) > 0
    assert cap.history()["Lp"].shape == (0,)
    assert float(opt.state[xi]["step"]) == 7
    t = cap.targeted_points(0)
    assert t.shape[1] == 2 and t.shape[0] == int(h["Num Targeted"][-1, 0])
    # a second capture continues from the optimiser's step count
    cap2 = P.Captured_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, opt, 1000, Seed=1, First_Epoch=7)
    cap2.run(2)
    assert float(opt.state[xi]["step"]) == 9
    with pytest.raises(TypeError):
        P.Captured_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, torch.optim.SGD([xi], lr=0.1), 100)
    with pytest.raises(ValueError):
        P.Captured_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, torch.optim.Adam([xi], lr=0.1), 100)
    with pytest.raises(ValueError):
        P.Captured_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W,
                          torch.optim.Adam([q for u in U for q in u.parameters()] + [xi], lr=0.1, weight_decay=0.1), 100)


def test_device_row_count_matches_sliced_launch():
    """pdl_launch_args.n_points_dev: a launch over a 3000-row buffer with the count 1777 on the device equals the launch over the
    first 1777 rows (loss, residual, gradients), bit for bit."""
    import pde_learn_b200 as P
    from pde_learn_b200 import engine
    dev, U, xi, mask, X, Y, B, derivs, lhs, rhs, W = _setup(P)
    pts = P.Generate_Points(B[0], 3000, dev, Seed=9)
    plan = engine.collocation_plan(U[0], derivs, lhs, rhs)
    prm = [q.detach() for q in engine.network_param_tensors(U[0])]
    md = engine.device_mask(mask, len(rhs), dev)
    outs = []
    for mode in (0, 1):
        n = 1777
        res = torch.zeros(3000, dtype=torch.float32, device=dev)
        st = torch.zeros(4, dtype=torch.float64, device=dev)
        gp = torch.zeros(plan.n_param_scalars, dtype=torch.float32, device=dev)
        gx = torch.zeros(len(rhs), dtype=torch.float32, device=dev)
        if mode == 0:
            plan.launch(pts[:n].contiguous(), prm, xi.detach(), md, None, 1.0 / n, True, res, None, gp, gx, st)
        else:
            cnt = torch.tensor([n], dtype=torch.int64, device=dev)
            plan.launch(pts, prm, xi.detach(), md, None, 123.0, True, res, None, gp, gx, st, n_points_dev=cnt)
        outs.append((res.clone(), st.clone(), gp.clone(), gx.clone()))
    a, b = outs
    assert torch.equal(a[0][:1777], b[0][:1777]) and float(b[0][1777:].abs().max()) == 0.0
    assert abs(float(a[1][0]) - float(b[1][0])) <= 1e-12 * abs(float(a[1][0]))
    assert float((a[2] - b[2]).abs().max()) <= 1e-6 * float(a[2].abs().max())
    assert float((a[3] - b[3]).abs().max()) <= 1e-6 * float(a[3].abs().max())
