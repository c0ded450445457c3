"""
Worker for tests/test_gpu_multi.py (launched with torch.distributed.run, one process per GPU, NCCL):
every rank trains on its shard of the same global collocation batch; rank 0 also repeats the step on the whole
batch without a process group and checks that losses, gradients-after-all-reduce and updated parameters agree.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pde_learn_b200 as P
from pde_learn_b200 import configs


def build(dev, seed):
    torch.manual_seed(seed)
    d, bounds, derivs, lhs, rhs = configs.config("kdv")
    U_List = [P.Network([d, 24, 24, 24, 1], "Rational", Device=dev) for _ in range(2)]
    K = len(rhs)
    g = torch.Generator().manual_seed(seed + 1)
    Xi = (0.1 * torch.randn(K, generator=g)).to(dev).requires_grad_(True)
    Mask = torch.zeros(K, dtype=torch.bool); Mask[5] = True
    return d, bounds, derivs, lhs, rhs, U_List, Xi, Mask


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d, bounds, derivs, lhs, rhs, U_List, Xi, Mask = build(dev, 3)
    N, ND = 4096 + 8 * world, 400
    W = {"Data": 1.0, "Coll": 1.0, "Lp": 1e-3, "L2": 1e-5}
    coll_all = [P.Generate_Points(bounds, N, dev, Seed=10 + s) for s in range(2)]
    x_all = [P.Generate_Points(bounds, ND, dev, Seed=20 + s) for s in range(2)]
    y_all = [torch.randn(ND, generator=torch.Generator().manual_seed(30 + s), dtype=torch.float64).to(dev) for s in range(2)]
    sh = lambda t, n: t[rank * (n // world):(rank + 1) * (n // world)].contiguous()
    # the shard sampled directly with First_Index equals the slice of the global batch
    mine = P.Generate_Points(bounds, N // world, dev, Seed=10, First_Index=rank * (N // world))
    assert torch.equal(mine, sh(coll_all[0], N))
    opt = torch.optim.Adam([q for U in U_List for q in U.parameters()] + [Xi], lr=1e-3)
    out = P.Training(U_List, Xi, Mask, [sh(c, N) for c in coll_all], [sh(x, ND) for x in x_all], [sh(y, ND) for y in y_all],
                     derivs, lhs, rhs, 0.1, W, opt, Process_Group=dist.group.WORLD, Residuals_Device="cuda")
    # replicas stay identical
    flat = torch.cat([q.detach().reshape(-1) for U in U_List for q in U.parameters()] + [Xi.detach()])
    ref = flat.clone(); dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), "replicas diverged"
    ok = True
    if rank == 0:
        d, bounds, derivs, lhs, rhs, V_List, Xi1, Mask1 = build(dev, 3)
        opt1 = torch.optim.Adam([q for U in V_List for q in U.parameters()] + [Xi1], lr=1e-3)
        one = P.Training(V_List, Xi1, Mask1, coll_all, x_all, y_all, derivs, lhs, rhs, 0.1, W, opt1, Residuals_Device="cuda")
        rel = lambda a, b: abs(a - b) / max(abs(b), 1e-30)
        errs = {k: max(rel(a, b) for a, b in zip(out[k], one[k])) for k in ("Coll Losses", "Data Losses", "L2 Losses", "Total Losses")}
        errs["Lp Loss"] = rel(out["Lp Loss"], one["Lp Loss"])
        flat1 = torch.cat([q.detach().reshape(-1) for U in V_List for q in U.parameters()] + [Xi1.detach()])
        errs["params_after_step_abs"] = float((flat - flat1).abs().max())
        errs["residual_shard"] = float((out["Residuals"][0] - one["Residuals"][0][:N // world]).abs().max())
        ok = all(v <= 2e-6 for k, v in errs.items() if k not in ("params_after_step_abs", "residual_shard"))
        ok = ok and errs["params_after_step_abs"] <= 2e-5 and errs["residual_shard"] == 0.0      # Adam step = 1e-3: 2 % of a step
        print(json.dumps({"world": world, "ok": ok, "errs": errs}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
