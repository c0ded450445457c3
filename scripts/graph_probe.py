"""A few captured epochs (for ncu): python scripts/graph_probe.py [workload] [N] [epochs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pde_learn_b200 as P
from pde_learn_b200 import configs

w = sys.argv[1] if len(sys.argv) > 1 else "heat"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
E = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
d, bounds, derivs, lhs, rhs = configs.config(w)
W = {"Data": 1.0, "Coll": 1.0, "Lp": 2e-4, "L2": 5e-6}
torch.manual_seed(0)
U = [P.Network([d] + [40] * 5 + [1], "Tanh", Device=dev)]
xi = torch.zeros(len(rhs), device=dev, requires_grad=True)
mask = torch.zeros(len(rhs), dtype=torch.bool)
X = [P.Generate_Points(bounds, 10_000, dev, Seed=77)]
Y = [torch.sin(X[0][:, 1].double()) * torch.exp(-0.1 * X[0][:, 0].double())]
opt = torch.optim.Adam(list(U[0].parameters()) + [xi], lr=1e-3)
cap = P.Captured_Epochs(U, xi, mask, X, Y, [bounds], derivs, lhs, rhs, 0.1, W, opt, N, Seed=1)
cap.run(E)
torch.cuda.synchronize()
print(cap.history()["Train Total"][:, 0])
