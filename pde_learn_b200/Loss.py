"""
Drop-in replacements for the reference loss functions (Code/Loss.py), same names, arguments and return
values, running in fused sm_100a kernels:

    Data_Loss(U, Inputs, Targets)                                          Loss.py:23-62
    Coll_Loss(U, Xi, Mask, Coll_Points, Derivatives, LHS_Term, RHS_Terms)  Loss.py:66-248
    Lp_Loss(Xi, Mask, p)                                                   Loss.py:252-312
    L2_Squared_Loss(U)                                                     Loss.py:316-361

Each returned loss is a torch scalar wired into autograd: the loss value AND all gradients (network
parameters, Rational coefficients, Xi) are produced by the same kernel launch; `.backward()` only scales
them by the upstream gradient.  Semantics kept from the reference: Mask[k] == True EXCLUDES term k and
leaves Xi.grad[k] exactly 0; Coll is a mean over all rows; Data is a float64 mean of
(float32 prediction - float64 target)^2; L^p uses detached IRLS weights with delta = 1e-7.

Extra keyword `Inv_N` (default 1/rows): when the rows are a shard of a larger batch, pass
1/global_rows so that summing the per-rank losses/gradients gives the global mean.
"""
from typing import List, Tuple

import torch

from . import engine as E


def Data_Loss(U, Inputs: torch.Tensor, Targets: torch.Tensor, Inv_N: float = None) -> torch.Tensor:
    plan = E.data_plan(U)
    X = E._check_cuda_f32(Inputs, "Inputs")
    Y = Targets
    if not Y.is_cuda:
        raise RuntimeError("Targets must be a CUDA tensor: this engine has no CPU path")
    Y = Y.reshape(-1).to(torch.float64).contiguous()          # reference data sets carry float64 targets
    assert Y.numel() == X.shape[0]
    params = [E._check_cuda_f32(p, "network parameter") for p in E.network_param_tensors(U)]
    want = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    inv_n = Inv_N if Inv_N is not None else 1.0 / max(int(X.shape[0]), 1)
    return E._DataLossFn.apply(plan, X.detach(), Y, inv_n, want, *params)


def Coll_Loss(U, Xi: torch.Tensor, Mask: torch.Tensor, Coll_Points: torch.Tensor, Derivatives: List,
              LHS_Term, RHS_Terms: List, Device: torch.device = None, Inv_N: float = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    assert torch.numel(Xi) == len(RHS_Terms)                   # Loss.py:124
    plan = E.collocation_plan(U, Derivatives, LHS_Term, RHS_Terms)
    P = E._check_cuda_f32(Coll_Points, "Coll_Points")
    assert P.dim() == 2 and P.shape[1] == plan.d, "Coll_Points must be [B, %d]" % plan.d
    xi = E._check_cuda_f32(Xi, "Xi")
    mask_dev = E.device_mask(Mask, plan.K, P.device)
    params = [E._check_cuda_f32(p, "network parameter") for p in E.network_param_tensors(U)]
    want = torch.is_grad_enabled() and (xi.requires_grad or any(p.requires_grad for p in params))
    inv_n = Inv_N if Inv_N is not None else 1.0 / max(int(P.shape[0]), 1)
    loss, residual = E._CollLossFn.apply(plan, P.detach(), mask_dev, inv_n, want, xi, *params)
    return loss, residual


def Lp_Loss(Xi: torch.Tensor, Mask: torch.Tensor, p: float) -> torch.Tensor:
    assert p > 0 and p < 2                                     # Loss.py:279
    xi = E._check_cuda_f32(Xi, "Xi")
    return E._LpLossFn.apply(xi, E.device_mask(Mask, int(xi.numel()), xi.device), float(p))


def L2_Squared_Loss(U) -> torch.Tensor:
    """Shape-[1] float32 tensor like the reference's (which lives on the CPU; ours stays on the device)."""
    params = [E._check_cuda_f32(p, "network parameter") for p in E.network_param_tensors(U)]
    return E._L2LossFn.apply(*params)
