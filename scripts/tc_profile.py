"""Cycle accounting of the tcgen05 forward engine (plans compiled with PDL_NVCC_FLAGS=-DPDL_TC_PROFILE): where the activation warps
and the issuer warp wait.  usage: PDL_NVCC_FLAGS=-DPDL_TC_PROFILE python scripts/tc_profile.py heat Tanh"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pde_learn_b200 as P
from pde_learn_b200 import configs, engine

w, act = sys.argv[1], sys.argv[2]
B = 1 << 20
dev = torch.device("cuda", 0)
d, bounds, derivs, lhs, rhs = configs.config(w)
torch.manual_seed(0)
U = P.Network([d] + [40] * 5 + [1], act, Device=dev)
K = len(rhs)
xi = (0.1 * torch.randn(K)).to(dev)
mask = torch.zeros(K, dtype=torch.uint8, device=dev)
pts = P.Generate_Points(bounds, B, dev, Seed=1)
plan = engine.collocation_plan(U, derivs, lhs, rhs)
prm = [q.detach() for q in engine.network_param_tensors(U)]
r = torch.zeros(B, device=dev); feats = torch.zeros((plan.J, B), device=dev); st = torch.zeros(4, dtype=torch.float64, device=dev)
for _ in range(3):
    plan.launch(pts, prm, xi, mask, None, 1.0 / B, False, r, feats, None, None, st)
torch.cuda.synchronize()
o = feats.reshape(-1)[:148 * 16].reshape(148, 16).cpu()
m = o.mean(0)
tiles = float(m[5])
print("%s %s: tiles/CTA %.1f" % (w, act, tiles))
for name, base in (("warpgroup 0 warp 0", 0), ("warpgroup 1 warp 4", 8)):
    t = float(m[base + 0])
    print("  %s: total %.0f cycles = %.0f per tile; wait ring slot %.1f%%, wait D (layer boundary) %.1f%%, wait tcgen05.ld %.1f%%, tile barrier %.1f%%"
          % (name, t, t / tiles, 100 * m[base + 1] / t, 100 * m[base + 2] / t, 100 * m[base + 3] / t, 100 * m[base + 4] / t))
t = float(m[6])
print("  issuer: total %.0f cycles; waiting for the ring (A operand) %.1f%%, issuing %.1f%% = %.1f cycles per MMA"
      % (t, 100 * m[7] / t, 100 * (1 - m[7] / t), (t - float(m[7])) / (tiles * (plan.J * 3 * 5 * 4))))
