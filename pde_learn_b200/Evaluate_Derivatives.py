"""
Derivative features D_alpha U by forward jet propagation.

The reference computes them one order at a time with torch.autograd.grad from a graph-attached tensor
(`Derivative_From_Derivative(Da, Db, Db_U, Coords)`, Code/Evaluate_Derivatives.py:20-207).  A fused kernel
cannot consume an autograd graph, so the entry point here takes the network instead of `Db_U`:

    Evaluate_Derivatives(U, Derivatives, Coords) -> [len(Derivatives), B] float32

Row j equals Derivative_From_Derivative(Derivatives[j], I, U(Coords), Coords) of the reference.
`Derivative_From_Derivative` itself is kept as a name that explains this and raises.
"""
from typing import List

import numpy as np
import torch

from . import engine as E
from .Derivative import Derivative
from .Term import Term


def Evaluate_Derivatives(U, Derivatives: List, Coords: torch.Tensor) -> torch.Tensor:
    # a plan whose "library" is just the requested derivatives: LHS = first one, RHS = the others
    lhs = Term([Derivatives[0]], [1])
    rhs = [Term([D], [1]) for D in Derivatives[1:]]
    plan = E.collocation_plan(U, Derivatives, lhs, rhs)
    P = E._check_cuda_f32(Coords, "Coords").detach()
    n = P.shape[0]
    dev = P.device
    feats = torch.empty((plan.J, n), dtype=torch.float32, device=dev)
    stats = torch.empty(4, dtype=torch.float64, device=dev)
    xi = torch.zeros(max(plan.K, 1), dtype=torch.float32, device=dev)
    mask = torch.zeros(max(plan.K, 1), dtype=torch.uint8, device=dev)
    params = [E._check_cuda_f32(p, "network parameter").detach() for p in E.network_param_tensors(U)]
    plan.launch(P, params, xi, mask, None, 1.0 / max(n, 1), False, None, feats, None, None, stats)
    rows = []
    for D in Derivatives:
        e = tuple(int(v) for v in np.asarray(D.Encoding).tolist())
        rows.append(plan.jets.index(e + (0,) * (4 - len(e))))
    return feats[rows]


def Derivative_From_Derivative(Da, Db, Db_U, Coords):
    raise NotImplementedError(
        "Derivative_From_Derivative differentiates an autograd graph (Code/Evaluate_Derivatives.py:104-205); the "
        "B200 engine propagates jets through the network instead. Call Evaluate_Derivatives(U, Derivatives, Coords).")


def Evaluate_Residual(U, Xi: torch.Tensor, Mask: torch.Tensor, Coords: torch.Tensor, Derivatives: List, LHS_Term, RHS_Terms: List
                      ) -> torch.Tensor:
    """
    PDE residual T_0(U) - sum_k xi_k T_k(U) at every row of Coords, forward-only kernel, any number of rows in one call.
    This is what Plot/Plot_One_Spatial_Dimension.py:149-178 gets from Coll_Loss in batches of 1000 points (pass the real
    mask here: the reference's plot passes an all-True mask and therefore shows b(U) only).
    """
    from .Loss import Coll_Loss
    with torch.no_grad():
        return Coll_Loss(U, Xi, Mask, Coords, Derivatives, LHS_Term, RHS_Terms)[1]


def Evaluate_Network(U, Coords: torch.Tensor) -> torch.Tensor:
    """U(Coords) as a [B] float32 tensor through the forward-only kernel (same numbers the losses see)."""
    return Evaluate_Derivatives(U, [Derivative(Encoding=np.array([0, 0]))], Coords)[0]
