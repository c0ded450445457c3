"""Kernel-only timing probe: fused fwd+adjoint launch per workload (run on a B200 via gpurun)."""
import ctypes as C
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pde_learn_b200 as P
from pde_learn_b200 import _lib as L
from pde_learn_b200 import configs, engine


def probe(w, act, B, bwd=True, reps=8):
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
    resid = torch.empty(B, dtype=torch.float32, device=dev)
    stats = torch.empty(4, dtype=torch.float64, device=dev)
    gp = torch.empty(plan.n_param_scalars, dtype=torch.float32, device=dev)
    gx = torch.empty(K, dtype=torch.float32, device=dev)
    ts = []
    flush = torch.empty(768 << 20, dtype=torch.uint8, device=dev) if os.environ.get("PDL_PROBE_FLUSH") else None
    for i in range(reps + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if flush is not None:
            flush.zero_()
        a.record()
        plan.launch(pts, prm, xi, mask, None, 1.0 / B, bwd, resid, None, gp if bwd else None, gx if bwd else None, stats)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    fl = plan.flops_per_point if bwd else plan.flops_per_point / 3
    return {"workload": w, "act": act, "B": B, "bwd": bwd, "ms": ms, "min_ms": min(ts), "pts_per_s": B / (ms * 1e-3),
            "tflops": fl * B / (ms * 1e-3) / 1e12, "warps": int(plan.info.warps_per_cta), "smem": int(plan.info.smem_bytes)}


if __name__ == "__main__":
    peak = C.c_double(0)
    L.check(L.lib().pdl_measure_fp32_peak(C.byref(peak), None))
    print(json.dumps({"fp32_peak_tflops": peak.value}))
    for w, act, B in [("heat", "Tanh", 1 << 20), ("heat", "Tanh", 1 << 16), ("heat", "Tanh", 10_000), ("burgers", "Tanh", 10_000),
                      ("kdv", "Tanh", 1 << 20), ("ks", "Tanh", 1 << 20), ("rd2d", "Tanh", 1 << 20), ("heat", "Rational", 1 << 20),
                      ("ks", "Rational", 1 << 20)]:
        print(json.dumps(probe(w, act, B)), flush=True)
    print(json.dumps(probe("heat", "Tanh", 1 << 20, bwd=False)), flush=True)
