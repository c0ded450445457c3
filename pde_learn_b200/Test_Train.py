"""
Training(...) / Testing(...): one epoch of the reference loop (Code/Test_Train.py:22-201, :205-330) with
the same arguments and the same returned dictionaries, every loss evaluated by the fused CUDA kernels.

What differs from the reference, on purpose:
  * no per-loss `.item()` host syncs inside the closure: loss scalars stay on the device and are read
    back with ONE copy after `Optimizer.step` (the values returned are those of the last closure call,
    as in the reference);
  * "Lp Loss" returned by Training is the true value (the reference returns 0.0 because its buffer is
    assigned without `nonlocal`, Test_Train.py:146; nobody reads it);
  * optional data parallelism: when `Process_Group` is given, the collocation rows (and data rows) this
    rank holds are a shard; losses and gradients are means over the GLOBAL row counts and the gradients
    of all parameters and Xi are all-reduced once per closure (NCCL on GPUs, gloo in CPU tests of the
    host logic).  L^p and L2 are replicated and enter each rank's total with weight 1/world_size.
"""
from typing import Dict, List

import torch

from .Loss import Coll_Loss, Data_Loss, L2_Squared_Loss, Lp_Loss
from .engine import closure_total


def _global_counts(local_counts: List[int], group) -> List[int]:
    if group is None:
        return list(local_counts)
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(local_counts, dtype=torch.int64, device=dev)
    dist.all_reduce(t, group=group)
    return [int(v) for v in t.tolist()]


def allreduce_gradients(tensors: List[torch.Tensor], group) -> None:
    """One flat all-reduce (sum) of the .grad of every tensor in `tensors`."""
    import torch.distributed as dist
    grads = [t.grad for t in tensors if t.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def Training(U_List: List, Xi: torch.Tensor, Mask: torch.Tensor, Coll_Points_List: List[torch.Tensor],
             Inputs_List: List[torch.Tensor], Targets_List: List[torch.Tensor], Derivatives: List, LHS_Term,
             RHS_Terms: List, p: float, Weights: Dict[str, float], Optimizer: torch.optim.Optimizer,
             Device: torch.device = None, Process_Group=None, Residuals_Device: str = "cpu") -> Dict:
    assert len(U_List) == len(Coll_Points_List)
    assert len(U_List) == len(Inputs_List)
    assert len(U_List) == len(Targets_List)
    S = len(U_List)
    for U in U_List:
        if not U.training:            # Module.train() walks every submodule; skip the walk when nothing changes
            U.train()
    world = 1
    if Process_Group is not None:
        import torch.distributed as dist
        world = dist.get_world_size(Process_Group)
    n_coll = _global_counts([int(P.shape[0]) for P in Coll_Points_List], Process_Group)
    n_data = _global_counts([int(X.shape[0]) for X in Inputs_List], Process_Group)
    last = {}

    def Closure() -> torch.Tensor:
        if torch.is_grad_enabled():
            Optimizer.zero_grad()
        # the whole closure is one autograd node (engine._ClosureFn): Lp + per network Coll/Data/L2 and the
        # weighted combination of their gradients are launched back to back, no per-loss host syncs
        total, scalars, residuals = closure_total(U_List, Xi, Mask, Coll_Points_List, Inputs_List, Targets_List, Derivatives,
                                                  LHS_Term, RHS_Terms, p, Weights,
                                                  inv_n_coll=[1.0 / max(n, 1) for n in n_coll],
                                                  inv_n_data=[1.0 / max(n, 1) for n in n_data], world=world, group=Process_Group)
        if total.requires_grad:
            total.backward()          # with a process group the closure node all-reduces its flat gradient buffer itself
        last.update(scalars=scalars, residuals=residuals)
        return total.reshape(1)

    Optimizer.step(Closure)

    # one device->host copy for all reported scalars: [coll_*, data_*, l2_*, total_*, lp]
    pack = last["scalars"].clone()
    if Process_Group is not None:
        import torch.distributed as dist
        pack[2 * S:3 * S] /= world                    # L2 and Lp are replicated, everything else is a partial sum
        pack[4 * S] /= world
        dist.all_reduce(pack, group=Process_Group)
    host = pack.cpu().tolist()
    res = [r.detach() for r in last["residuals"]]
    if Residuals_Device == "cpu":
        res = [r.cpu() for r in res]
    return {"Residuals": res,
            "Coll Losses": host[0:S],
            "Data Losses": host[S:2 * S],
            "Lp Loss": host[4 * S],
            "L2 Losses": host[2 * S:3 * S],
            "Total Losses": host[3 * S:4 * S]}


def Testing(U_List: List, Xi: torch.Tensor, Mask: torch.Tensor, Coll_Points_List: List[torch.Tensor],
            Inputs_List: List[torch.Tensor], Targets_List: List[torch.Tensor], Derivatives: List, LHS_Term,
            RHS_Terms: List, p: float, Weights: Dict[str, float], Device: torch.device = None) -> Dict:
    assert len(U_List) == len(Coll_Points_List)
    assert len(U_List) == len(Inputs_List)
    assert len(U_List) == len(Targets_List)
    S = len(U_List)
    for U in U_List:
        if U.training:
            U.eval()
    with torch.no_grad():
        vals = [Lp_Loss(Xi=Xi, Mask=Mask, p=p).to(torch.float64)]
        for i in range(S):
            vals.append(Data_Loss(U=U_List[i], Inputs=Inputs_List[i], Targets=Targets_List[i]).to(torch.float64))
            vals.append(Coll_Loss(U=U_List[i], Xi=Xi, Mask=Mask, Coll_Points=Coll_Points_List[i], Derivatives=Derivatives,
                                  LHS_Term=LHS_Term, RHS_Terms=RHS_Terms)[0].to(torch.float64))
            vals.append(L2_Squared_Loss(U=U_List[i]).reshape(()).to(torch.float64))
        host = torch.stack(vals).cpu().tolist()
    lp = host[0]
    data = [host[1 + 3 * i] for i in range(S)]
    coll = [host[2 + 3 * i] for i in range(S)]
    l2 = [host[3 + 3 * i] for i in range(S)]
    total = [Weights["Data"] * data[i] + Weights["Coll"] * coll[i] + Weights["Lp"] * lp + Weights["L2"] * l2[i]
             for i in range(S)]
    return {"Data Losses": data, "Coll Losses": coll, "Lp Loss": lp, "L2 Losses": l2, "Total Losses": total}
