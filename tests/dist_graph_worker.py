"""
Worker for tests/test_gpu_multi.py::test_captured_epochs_data_parallel (torch.distributed.run, one process per GPU, NCCL):
the whole {sample -> loss -> all-reduce -> Adam -> targeted selection} epoch captured as ONE CUDA graph per rank, with the two
NCCL all-reduces inside the graph.  Every rank runs Run_Epochs(..., Process_Group, Use_CUDA_Graph=True) on its
shard; rank 0 repeats the run in one process (no group, same seeds, all rows) and the histories / final parameters must agree.
With --time it also reports epochs/s of the eager and the captured data-parallel loop on config 1 (Burgers, 10 000 rows).
"""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import pde_learn_b200 as P
from pde_learn_b200 import configs


def build(dev, widths, act, seed=0, S=1):
    d, bounds, derivs, lhs, rhs = configs.config("burgers")
    torch.manual_seed(seed)
    U = [P.Network([d] + widths + [1], act, Device=dev) for _ in range(S)]
    xi = (0.05 * torch.randn(len(rhs))).to(dev).requires_grad_(True)
    mask = torch.zeros(len(rhs), dtype=torch.bool); mask[3] = True
    g = torch.Generator().manual_seed(5)
    lo = torch.tensor(bounds[:, 0], dtype=torch.float32); hi = torch.tensor(bounds[:, 1], dtype=torch.float32)
    X = [(lo + (hi - lo) * torch.rand(320, d, generator=g)).to(dev) for _ in range(S)]
    Y = [torch.sin(x[:, 1].double()) * torch.exp(-0.1 * x[:, 0].double()) for x in X]
    W = {"Data": 1.0, "Coll": 1.0, "Lp": 2e-4, "L2": 5e-6}
    return U, xi, mask, X, Y, [bounds] * S, derivs, lhs, rhs, W


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
    E, N = 10, 4096                                            # rows in total, for any world size
    U, xi, mask, X, Y, B, derivs, lhs, rhs, W = build(dev, [20, 20, 20], "Rational", S=2)
    U1 = copy.deepcopy(U); xi1 = xi.detach().clone().requires_grad_(True)
    opt = torch.optim.Adam([q for u in U for q in u.parameters()] + [xi], lr=2e-3)
    h = P.Run_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, opt, E, N, Seed=3, Process_Group=pg, Use_CUDA_Graph=True)
    nt = torch.tensor(h["Num Targeted"], device=dev); dist.all_reduce(nt)
    flat = torch.cat([q.detach().reshape(-1) for u in U for q in u.parameters()] + [xi.detach()])
    ref = flat.clone(); dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), "replicas diverged"
    ok, out = True, {"world": world}
    if rank == 0:
        opt1 = torch.optim.Adam([q for u in U1 for q in u.parameters()] + [xi1], lr=2e-3)
        h1 = P.Run_Epochs(U1, xi1, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, opt1, E, N, Seed=3, Use_CUDA_Graph=True)
        errs = {}
        for key in ("Train Coll", "Train Data", "Train Total", "L2"):
            errs[key] = float(np.max(np.abs(h[key] - h1[key]) / np.maximum(np.abs(h1[key]), 1e-30)))
        errs["Lp"] = float(np.max(np.abs(h["Lp"] - h1["Lp"]) / np.abs(h1["Lp"])))
        errs["first_epochs"] = float(np.max(np.abs(h["Train Total"][:2] - h1["Train Total"][:2]) / np.abs(h1["Train Total"][:2])))
        errs["num_targeted_diff"] = int(np.abs(nt.cpu().numpy() - h1["Num Targeted"]).max())
        flat1 = torch.cat([q.detach().reshape(-1) for u in U1 for q in u.parameters()] + [xi1.detach()])
        errs["params_abs"] = float((flat - flat1).abs().max())
        ok = (errs["first_epochs"] <= 2e-5 and all(errs[k] <= 1e-3 for k in ("Train Coll", "Train Data", "Train Total", "L2")) and
              errs["Lp"] <= 3e-2 and errs["num_targeted_diff"] <= 2 and errs["params_abs"] <= 5e-4 and int(h1["Num Targeted"].max()) > 0)
        out.update(errs=errs, targeted_max=int(h1["Num Targeted"].max()))
    if "--time" in sys.argv:
        # config 1 (Burgers, 5 x 40 tanh, 10 000 rows in total): epochs/s of the eager and of the captured data-parallel loop
        n_tot = 10000 - 10000 % world
        rates = {}
        for mode in ("eager", "graph"):
            U, xi, mask, X, Y, B, derivs, lhs, rhs, W = build(dev, [40] * 5, "Tanh")
            opt = torch.optim.Adam([q for u in U for q in u.parameters()] + [xi], lr=1e-3)
            kw = dict(Seed=7, Process_Group=pg, Use_CUDA_Graph=(mode == "graph"))
            P.Run_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, opt, 20, n_tot, **kw)          # warm-up (plans, capture)
            if mode == "graph":
                cap = P.Captured_Epochs(U, xi, mask, [x[rank * (320 // world):(rank + 1) * (320 // world)] for x in X],
                                        [y[rank * (320 // world):(rank + 1) * (320 // world)] for y in Y], B, derivs, lhs, rhs, 0.1, W, opt,
                                        n_tot, Seed=8, Process_Group=pg).capture()
                cap.run(10)
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode == "graph":
                cap.run(200); cap.history()
            else:
                P.Run_Epochs(U, xi, mask, X, Y, B, derivs, lhs, rhs, 0.1, W, opt, 200, n_tot, **kw)
            torch.cuda.synchronize(); dist.barrier()
            rates[mode] = 200 / (time.perf_counter() - t0)
            if mode == "graph":
                cap.close()                      # a live CUDA graph with captured NCCL kernels blocks destroy_process_group()
        out["config1_epochs_per_s"] = rates
        out["config1_graph_speedup"] = rates["graph"] / rates["eager"]
    if rank == 0:
        out["ok"] = ok
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
