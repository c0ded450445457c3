"""
Generates tests/golden/burgers_run.npz: a short, fully seeded training run of the UNMODIFIED reference on real
data (Matlab/Data/Burgers_Sine.mat, noisy subsample as in Data/From_MATLAB.py:75-148), i.e. the epoch loop of
Code/main.py:281-384 (random collocation points + targeted points, Training with Adam) on the CPU in float32.

The GPU test replays the same epochs through pde_learn_b200 (same initial weights, same data, same per-epoch
random points) and checks the Xi trajectory and the thresholded support (|xi_k| >= 5e-4, main.py:31).

    python tests/golden/make_burgers_run.py
"""
import json
import os
import sys

import numpy as np
import scipy.io
import torch

REF = "/root/reference/Code"
for p in (REF, os.path.join(REF, "Classes"), os.path.join(REF, "Readers")):
    sys.path.insert(0, p)
HERE = os.path.dirname(os.path.abspath(__file__))

from Network import Network                                      # noqa: E402  (reference)
from Test_Train import Training                                   # noqa: E402
from Library_Reader import Read_Library                          # noqa: E402

EPOCHS, N_COLL, N_DATA, NOISE, LR = 240, 3000, 2000, 0.25, 1e-3
WIDTHS = [2, 20, 20, 20, 20, 20, 1]
WEIGHTS = {"Data": 1.0, "Coll": 1.0, "Lp": 0.0, "L2": 5e-6}          # burn-in stage of the reference workflow (Settings.txt:37)
P_EXP = 0.1
CHECK_EVERY = 40


def epoch_points(bounds, e):
    """The epoch's random collocation points: seeded torch CPU generator (identical on every machine)."""
    g = torch.Generator().manual_seed(1000 + e)
    u = torch.rand(N_COLL, 2, generator=g, dtype=torch.float32)
    lo = torch.tensor(bounds[:, 0], dtype=torch.float32); hi = torch.tensor(bounds[:, 1], dtype=torch.float32)
    return lo + (hi - lo) * u


if __name__ == "__main__":
    torch.set_num_threads(8)
    m = scipy.io.loadmat("/root/reference/Matlab/Data/Burgers_Sine.mat")
    t = m["t"].reshape(-1).astype(np.float32); x = m["x"].reshape(-1).astype(np.float32)
    usol = np.real(m["usol"]).astype(np.float32)
    rs = np.random.RandomState(0)
    noisy = usol + NOISE * np.std(usol) * rs.randn(*usol.shape)            # float64, like From_MATLAB.py:92
    tt, xx = np.meshgrid(t, x)
    coords = np.hstack((tt.flatten().reshape(-1, 1), xx.flatten().reshape(-1, 1)))
    idx = rs.choice(coords.shape[0], N_DATA, replace=False)
    inputs = torch.from_numpy(coords[idx].astype(np.float32))
    targets = torch.from_numpy(noisy.flatten()[idx])                       # float64 targets
    bounds = np.array([[t[0], t[-1]], [x[0], x[-1]]], dtype=np.float32)

    Derivs, LHS, RHS = Read_Library("/root/reference/Library.txt")
    K = len(RHS)
    torch.manual_seed(0)
    U = Network(Widths=WIDTHS, Hidden_Activation="Tanh", Output_Activation="None")
    p0 = [q.detach().numpy().copy() for q in U.parameters()]
    Xi = torch.zeros(K, dtype=torch.float32, requires_grad=True)
    Mask = torch.zeros(K, dtype=torch.bool)
    Opt = torch.optim.Adam(list(U.parameters()) + [Xi], lr=LR)

    targeted = torch.empty((0, 2), dtype=torch.float32)
    xi_hist, n_targeted, coll_hist, data_hist = [], [], [], []
    for e in range(EPOCHS):
        pts = torch.vstack((epoch_points(bounds, e), targeted))
        out = Training([U], Xi, Mask, [pts], [inputs], [targets], Derivs, LHS, RHS, P_EXP, WEIGHTS, Opt)
        a = torch.abs(out["Residuals"][0]); head = a[:N_COLL]
        cutoff = head.mean() + 3 * head.std()
        targeted = pts[torch.greater_equal(a, cutoff), :].detach()
        n_targeted.append(int(targeted.shape[0])); coll_hist.append(out["Coll Losses"][0]); data_hist.append(out["Data Losses"][0])
        if (e + 1) % CHECK_EVERY == 0:
            xi_hist.append(Xi.detach().numpy().copy())
            print("epoch %4d coll %.4e data %.4e targeted %d  |xi|max %.4f" % (e + 1, coll_hist[-1], data_hist[-1], n_targeted[-1], float(Xi.abs().max())))

    out = {"inputs": inputs.numpy(), "targets": targets.numpy(), "bounds": bounds, "widths": np.array(WIDTHS),
           "epochs": EPOCHS, "n_coll": N_COLL, "lr": LR, "p": P_EXP, "weights": np.array(json.dumps(WEIGHTS)),
           "check_every": CHECK_EVERY, "xi_hist": np.stack(xi_hist), "n_targeted": np.array(n_targeted),
           "coll_hist": np.array(coll_hist), "data_hist": np.array(data_hist), "n_params": len(p0)}
    for i, a in enumerate(p0):
        out["p0_%02d" % i] = a
    for i, q in enumerate(U.parameters()):
        out["p1_%02d" % i] = q.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "burgers_run.npz"), **out)
    sup = np.abs(xi_hist[-1]) >= 5e-4
    print("final support:", np.nonzero(sup)[0].tolist(), "xi:", np.round(xi_hist[-1], 4).tolist())
