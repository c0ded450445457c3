"""
Generates the *_trained.npz golden cases: the same float64 reference outputs as make_golden.py, but for networks and coefficient
vectors taken from the END of a training run of the unmodified reference (Code/Test_Train.py:22 Training with Adam on real data,
Matlab/Data/*.mat subsampled with noise as Data/From_MATLAB.py does) instead of the Xavier initialisation: larger ||W||, saturated
tanh units, non-trivial Xi -- the regime where the split-precision contractions (bf16 cross terms) have the least headroom.

    python tests/golden/make_golden_trained.py          (a few minutes on 8 cores)
"""
import os
import sys

import numpy as np
import scipy.io
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G                                          # noqa: E402  (puts the reference on sys.path)
from make_golden import Network, Training, config_library       # noqa: E402


def load_mat(name, n_data, noise, seed):
    m = scipy.io.loadmat("/root/reference/Matlab/Data/%s.mat" % name)
    t = m["t"].reshape(-1).astype(np.float32); x = m["x"].reshape(-1).astype(np.float32)
    usol = np.real(m["usol"]).astype(np.float32)
    rs = np.random.RandomState(seed)
    noisy = usol + noise * np.std(usol) * rs.randn(*usol.shape)
    tt, xx = np.meshgrid(t, x)
    coords = np.hstack((tt.flatten().reshape(-1, 1), xx.flatten().reshape(-1, 1)))
    idx = rs.choice(coords.shape[0], n_data, replace=False)
    bounds = [(float(t[0]), float(t[-1])), (float(x[0]), float(x[-1]))]
    return torch.from_numpy(coords[idx].astype(np.float32)), torch.from_numpy(noisy.flatten()[idx]), bounds


def train(name, cfg, mat, hidden, act, epochs, n_coll, lr, seed, n_pts, masked=()):
    d, lhs, rhs, _ = config_library(cfg)
    inputs, targets, bounds = load_mat(mat, 2000, 0.1, seed)
    Derivs, LHS, RHS, encs = G.ref_terms(lhs, rhs)
    K = len(rhs)
    torch.manual_seed(seed)
    U = Network(Widths=[d] + hidden + [1], Hidden_Activation=act, Output_Activation="None")
    w0 = max(float(L.weight.abs().max()) for L in U.Layers)
    Xi = torch.zeros(K, dtype=torch.float32, requires_grad=True)
    Mask = torch.zeros(K, dtype=torch.bool)
    for k in masked:
        Mask[k] = True
    Opt = torch.optim.Adam(list(U.parameters()) + [Xi], lr=lr)
    W = {"Data": 1.0, "Coll": 1.0, "Lp": 0.0, "L2": 0.000005}          # the burn-in stage of the reference workflow (Settings.txt:37)
    g = torch.Generator().manual_seed(seed + 100)
    for e in range(epochs):
        pts = G.points_in(bounds, n_coll, g)
        out = Training([U], Xi, Mask, [pts], [inputs], [targets], Derivs, LHS, RHS, 0.1, W, Opt)
        if (e + 1) % 100 == 0:
            print("  %s epoch %4d coll %.3e data %.3e |xi|max %.4f" % (name, e + 1, out["Coll Losses"][0], out["Data Losses"][0],
                                                                      float(Xi.abs().max())), flush=True)
    w1 = max(float(L.weight.abs().max()) for L in U.Layers)
    print("  %s: max|W| %.3f -> %.3f after %d epochs" % (name, w0, w1, epochs))
    for q in U.parameters():
        q.requires_grad_(True)
    G.run_case(name, d, lhs, rhs, bounds, hidden, act, n_pts, seed + 7, masked=masked, U_trained=U, xi_trained=Xi)


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["burgers", "ks", "ks_rat"]
    if "burgers" in which:        # the architecture / data of make_burgers_run.py (5 x 20 tanh, shipped Library.txt), 300 epochs
        train("burgers_tanh20_trained", "burgers", "Burgers_Sine", [20] * 5, "Tanh", 300, 3000, 1e-3, 31, 128)
    if "ks" in which:             # KS library (4th-order jets, 30 terms) on a 5 x 40 tanh network after 500 Adam epochs
        train("ks_tanh_trained", "ks", "KS_Sine", [40] * 5, "Tanh", 500, 1000, 2e-3, 32, 96)
    if "ks_rat" in which:         # the reference's default activation, 500 epochs, one masked term
        train("ks_rational_trained", "ks", "KS_Sine", [40] * 5, "Rational", 500, 1000, 1e-3, 33, 64, masked=(7,))
