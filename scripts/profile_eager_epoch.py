"""cProfile of the eager epoch loop at a small batch (where host overhead dominates): python scripts/profile_eager_epoch.py"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pde_learn_b200 as P
from pde_learn_b200 import configs

dev = torch.device("cuda", 0)
d, bounds, derivs, lhs, rhs = configs.config("burgers")
W = {"Data": 1.0, "Coll": 1.0, "Lp": 2e-4, "L2": 5e-6}
torch.manual_seed(0)
U = [P.Network([d] + [40] * 5 + [1], "Tanh", Device=dev)]
xi = torch.zeros(len(rhs), device=dev, requires_grad=True)
mask = torch.zeros(len(rhs), dtype=torch.bool)
X = [P.Generate_Points(bounds, 10_000, dev, Seed=77)]
Y = [torch.sin(X[0][:, 1].double())]
opt = torch.optim.Adam(list(U[0].parameters()) + [xi], lr=1e-3)
P.Run_Epochs(U, xi, mask, X, Y, [bounds], derivs, lhs, rhs, 0.1, W, opt, 20, 10_000, Seed=1)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
P.Run_Epochs(U, xi, mask, X, Y, [bounds], derivs, lhs, rhs, 0.1, W, opt, 300, 10_000, Seed=2)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
