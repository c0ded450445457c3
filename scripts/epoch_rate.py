"""Whole-epoch rate (sample -> closure -> Adam -> targeted selection) at small batches: eager Run_Epochs vs Captured_Epochs.
usage: python scripts/epoch_rate.py [workload=burgers] [N=10000] [epochs=200]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pde_learn_b200 as P
from pde_learn_b200 import configs

w = sys.argv[1] if len(sys.argv) > 1 else "burgers"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
E = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda", 0)
d, bounds, derivs, lhs, rhs = configs.config(w)
W = {"Data": 1.0, "Coll": 1.0, "Lp": 2e-4, "L2": 5e-6}
out = {"workload": w, "rows": N, "epochs": E}
for mode in ("eager", "graph"):
    torch.manual_seed(0)
    U = [P.Network([d] + [40] * 5 + [1], "Tanh", Device=dev)]
    xi = torch.zeros(len(rhs), device=dev, requires_grad=True)
    mask = torch.zeros(len(rhs), dtype=torch.bool)
    X = [P.Generate_Points(bounds, 10_000, dev, Seed=77)]
    Y = [torch.sin(X[0][:, 1].double()) * torch.exp(-0.1 * X[0][:, 0].double())]
    opt = torch.optim.Adam(list(U[0].parameters()) + [xi], lr=1e-3)
    if mode == "eager":
        P.Run_Epochs(U, xi, mask, X, Y, [bounds], derivs, lhs, rhs, 0.1, W, opt, 10, N, Seed=1)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = P.Run_Epochs(U, xi, mask, X, Y, [bounds], derivs, lhs, rhs, 0.1, W, opt, E, N, Seed=2)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        extra = {}
    else:
        t0 = time.perf_counter()
        cap = P.Captured_Epochs(U, xi, mask, X, Y, [bounds], derivs, lhs, rhs, 0.1, W, opt, N, Seed=2).capture()
        torch.cuda.synchronize()
        t_cap = time.perf_counter() - t0
        cap.run(10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cap.run(E)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h = cap.history()
        extra = {"capture_ms": 1e3 * t_cap}
    out[mode] = dict({"ms_per_epoch": 1e3 * dt / E, "epochs_per_s": E / dt, "points_per_s": N * E / dt, "final_total": float(h["Train Total"][-1, 0]),
                      "targeted_last": int(h["Num Targeted"][-1, 0])}, **extra)
out["speedup"] = out["eager"]["ms_per_epoch"] / out["graph"]["ms_per_epoch"]
print(json.dumps(out))
