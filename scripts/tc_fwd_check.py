"""tcgen05 forward engine check (run on a B200): residual / loss of the forward-only launch against the fused fwd+adjoint launch
(mma.sync engine) on all rows, features and residual against the float64 oracle on a subset, and timings."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pde_learn_b200 as P
from pde_learn_b200 import configs, engine
from oracle import pde_oracle as O          # checker only (diagnostic script)


def run(w, act, B, n_oracle=1500):
    dev = torch.device("cuda", 0)
    d, bounds, derivs, lhs, rhs = configs.config(w)
    torch.manual_seed(0)
    U = P.Network([d] + [40] * 5 + [1], act, Device=dev)
    with torch.no_grad():
        for l in U.Layers:
            l.bias.add_(0.1 * torch.randn_like(l.bias))
    K = len(rhs)
    xi = (0.1 * torch.randn(K)).to(dev)
    mask = torch.zeros(K, dtype=torch.uint8, device=dev); mask[1] = 1
    pts = P.Generate_Points(bounds, B, dev, Seed=1)
    plan = engine.collocation_plan(U, derivs, lhs, rhs)
    prm = [q.detach() for q in engine.network_param_tensors(U)]
    r_f = torch.zeros(B, device=dev); r_b = torch.zeros(B, device=dev)
    feats = torch.zeros((plan.J, B), device=dev)
    st_f = torch.zeros(4, dtype=torch.float64, device=dev); st_b = torch.zeros(4, dtype=torch.float64, device=dev)
    gp = torch.empty(plan.n_param_scalars, device=dev); gx = torch.empty(K, device=dev)
    plan.launch(pts, prm, xi, mask, None, 1.0 / B, False, r_f, feats, None, None, st_f)
    plan.launch(pts, prm, xi, mask, None, 1.0 / B, True, r_b, None, gp, gx, st_b)
    torch.cuda.synchronize()
    out = {"workload": w, "act": act, "B": B, "J": plan.J,
           "resid_fwd_vs_bwd_relmax": float((r_f - r_b).abs().max() / r_b.abs().max()),
           "loss_fwd": float(st_f[0]), "loss_bwd": float(st_b[0]), "loss_rel": abs(float(st_f[0]) - float(st_b[0])) / abs(float(st_b[0])),
           "sum_abs_rel": abs(float(st_f[2]) - float(st_b[2])) / abs(float(st_b[2]))}
    # oracle on a subset (float64 nested autograd)
    n = min(n_oracle, B)
    idx = torch.linspace(0, B - 1, n).long().to(dev)
    nl = len(U.Layers)
    Ws = [l.weight.detach().cpu() for l in U.Layers]; bs = [l.bias.detach().cpu() for l in U.Layers]
    ra = rb = []
    if act == "Rational":
        ra = [a.a.detach().cpu() for a in list(U.Activation_Functions)[:-1]]; rb = [a.b.detach().cpu() for a in list(U.Activation_Functions)[:-1]]
    net = O.NetParams(Ws, bs, act.lower(), list(ra), list(rb)).to(torch.float64, False)
    od, olhs, orhs, _ = O.config_library(w)
    X = pts[idx].cpu().double().requires_grad_(True)
    fe = O.derivative_features(lambda Z: O.mlp_forward(net, Z), X, [j[:od] for j in plan.jets])
    _l, orr, _ = O.coll_loss(lambda Z: O.mlp_forward(net, Z), xi.cpu().double(), [bool(m) for m in mask.cpu()], pts[idx].cpu().double(), olhs, orhs)
    errs = []
    for c, j in enumerate(plan.jets):
        ref = fe[tuple(j[:od])].detach()
        errs.append(float((feats[c][idx].cpu().double() - ref).abs().max() / ref.abs().max()))
    out["feature_relmax_vs_oracle"] = errs
    out["resid_relmax_vs_oracle"] = float((r_f[idx].cpu().double() - orr.detach()).abs().max() / orr.detach().abs().max())
    out["resid_bwd_relmax_vs_oracle"] = float((r_b[idx].cpu().double() - orr.detach()).abs().max() / orr.detach().abs().max())
    # timing of the forward-only launch
    flush = torch.empty(768 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(11):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        a.record()
        plan.launch(pts, prm, xi, mask, None, 1.0 / B, False, r_f, None, None, None, st_f)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    out["fwd_ms"] = statistics.median(ts); out["fwd_min_ms"] = min(ts)
    out["fwd_tc"] = os.environ.get("PDL_FWD_TC", "default")
    return out


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20) + 77
    cases = [a.split(":") for a in sys.argv[2:]] or [["heat", "Tanh"], ["heat", "Rational"], ["burgers", "Tanh"], ["kdv", "Rational"]]
    for w, act in cases:
        print(json.dumps(run(w, act, B)), flush=True)
