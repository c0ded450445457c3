"""
Host glue between torch tensors and the C ABI (libpdelearn_b200.so): plan cache, workspaces, launches and
the torch.autograd.Function wrappers.  Everything numerical happens in the CUDA kernels; this file only
marshals pointers and hands gradients to autograd.  There is no CPU path: tensors must live on a CUDA
device and the native library must be built, otherwise the calls raise.
"""
import ctypes as C
import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

_RATIONAL_NAMES = ("Rational",)


# ----------------------------------------------------------------------------------------------
# network introspection (works for pde_learn_b200.Network and for the reference's Network class)
# ----------------------------------------------------------------------------------------------

def network_signature(U) -> Tuple[Tuple[int, ...], int]:
    """(widths, activation id) of a Network-like module; raises for anything the kernels cannot run."""
    if not hasattr(U, "Layers") or not hasattr(U, "Activation_Functions"):
        raise TypeError("U must be a Network (Layers / Activation_Functions). Arbitrary callables would need "
                        "autograd on the CPU; this engine has no such fallback.")
    layers = list(U.Layers)
    widths = [int(layers[0].weight.shape[1])] + [int(l.weight.shape[0]) for l in layers]
    acts = list(U.Activation_Functions)
    if type(acts[-1]).__name__ != "Identity":
        raise ValueError("output activation must be None/Identity (Code/main.py:139)")
    names = {type(a).__name__ for a in acts[:-1]}
    if len(names) != 1:
        raise ValueError("hidden layers must share one activation type, got %s" % sorted(names))
    name = names.pop()
    if name == "Tanh":
        act = L.PDL_ACT_TANH
    elif name in _RATIONAL_NAMES:
        act = L.PDL_ACT_RATIONAL
    else:
        raise ValueError("activation %s is not implemented in the CUDA kernels (Tanh, Rational); no fallback" % name)
    return tuple(widths), act


def network_param_tensors(U) -> List[torch.Tensor]:
    """Parameter tensors in the order of list(U.parameters()) (SURVEY 'Parameter order')."""
    out: List[torch.Tensor] = []
    for l in U.Layers:
        out += [l.weight, l.bias]
    for a in list(U.Activation_Functions)[:-1]:
        if type(a).__name__ in _RATIONAL_NAMES:
            out += [a.a, a.b]
    return out


def _check_cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: this engine has no CPU path" % what)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (what, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------
# plans
# ----------------------------------------------------------------------------------------------

class Plan:
    """One compiled kernel family: architecture + library of terms (or the data-loss residual)."""

    def __init__(self, d, widths, activation, mode, encodings, terms):
        self.desc, self._keep = L.make_desc(d, widths, activation, mode, encodings, terms)
        h = C.c_void_p()
        L.check(L.lib().pdl_plan_create(C.byref(self.desc), C.byref(h)))
        self.handle = h
        info = L.PlanInfo()
        L.check(L.lib().pdl_plan_info_get(h, C.byref(info)))
        self.info = info
        self.d, self.J, self.K = int(info.d), int(info.n_jets), int(info.n_rhs_terms)
        self.n_param_tensors, self.n_param_scalars = int(info.n_param_tensors), int(info.n_param_scalars)
        self.flops_per_point = int(info.flops_per_point)
        self.jets = [tuple(int(info.jet_encodings[c][a]) for a in range(4)) for c in range(self.J)]
        self._ws: Dict[int, torch.Tensor] = {}
        self._sm: Dict[int, int] = {}

    def workspace(self, device: torch.device) -> Tuple[torch.Tensor, int]:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._ws:
            with torch.cuda.device(idx):
                n = C.c_int32(0)
                L.check(L.lib().pdl_device_sm_count(C.byref(n)))
                self._sm[idx] = int(n.value)
                nbytes = int(L.lib().pdl_plan_workspace_bytes(self.handle, self._sm[idx]))
                self._ws[idx] = torch.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", idx))
        return self._ws[idx], self._sm[idx]

    def launch(self, points, params, xi, mask, targets, inv_n, backward, residual, features, grad_params, grad_xi,
               stats, n_points_dev=None):
        dev = points.device
        ws, n_sm = self.workspace(dev)
        arr = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        a = L.LaunchArgs()
        a.points = points.data_ptr(); a.n_points = int(points.shape[0])
        a.params = arr; a.n_params = len(params)
        a.xi = xi.data_ptr() if xi is not None else None
        a.mask = mask.data_ptr() if mask is not None else None
        a.targets = targets.data_ptr() if targets is not None else None
        a.inv_n = float(inv_n); a.backward = 1 if backward else 0
        a.residual = residual.data_ptr() if residual is not None else None
        a.features = features.data_ptr() if features is not None else None
        a.grad_params = grad_params.data_ptr() if grad_params is not None else None
        a.grad_xi = grad_xi.data_ptr() if grad_xi is not None else None
        a.stats = stats.data_ptr()
        a.workspace = ws.data_ptr(); a.workspace_bytes = ws.numel(); a.n_ctas = n_sm
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        a.n_points_dev = n_points_dev.data_ptr() if n_points_dev is not None else None   # device int64[1]: rows <= points.shape[0]
        with torch.cuda.device(dev):
            L.check(L.lib().pdl_loss_launch(self.handle, C.byref(a)))


_PLANS: Dict[tuple, Plan] = {}


def _enc_tuple(D) -> Tuple[int, ...]:
    return tuple(int(v) for v in np.asarray(D.Encoding).tolist())


_KEY_CACHE: Dict[tuple, tuple] = {}


def library_key(Derivatives, LHS_Term, RHS_Terms):
    """Hashable description of the library: distinct encodings + (encoding index, power) lists, LHS first.
    Memoised on the identity of the Derivative/Term objects (the cache entry keeps them alive, so an id cannot be reused):
    the epoch loop passes the same objects every call and rebuilding the key costs ~0.1 ms of host time per closure."""
    # + a cheap fingerprint of the mutable parts: Term.Append (public API kept from the reference) changes a term in place
    ident = (tuple(map(id, Derivatives)), id(LHS_Term), tuple(map(id, RHS_Terms)),
             tuple((T.Num_Sub_Terms, tuple(T.Powers)) for T in (LHS_Term, *RHS_Terms)))
    hit = _KEY_CACHE.get(ident)
    if hit is not None:
        return hit[0]
    key = _library_key(Derivatives, LHS_Term, RHS_Terms)
    if len(_KEY_CACHE) > 64:
        _KEY_CACHE.clear()
    _KEY_CACHE[ident] = (key, list(Derivatives), LHS_Term, list(RHS_Terms))
    return key


def _library_key(Derivatives, LHS_Term, RHS_Terms):
    encs: List[Tuple[int, ...]] = []

    def idx(D):
        e = _enc_tuple(D)
        e = e + (0,) * (4 - len(e))
        if e not in encs:
            encs.append(e)
        return encs.index(e)

    for D in Derivatives:
        idx(D)
    terms = []
    for T in [LHS_Term] + list(RHS_Terms):
        terms.append(tuple((idx(D), int(p)) for D, p in zip(T.Derivatives, T.Powers)))
    return tuple(encs), tuple(terms)


def collocation_plan(U, Derivatives, LHS_Term, RHS_Terms) -> Plan:
    widths, act = network_signature(U)
    encs, terms = library_key(Derivatives, LHS_Term, RHS_Terms)
    key = (widths, act, L.PDL_MODE_COLLOCATION, encs, terms)
    if key not in _PLANS:
        _PLANS[key] = Plan(widths[0], widths, act, L.PDL_MODE_COLLOCATION, encs, [list(t) for t in terms])
    return _PLANS[key]


def validate_configuration(d: int, Hidden_Widths, Hidden_Activation: str, Derivatives, LHS_Term, RHS_Terms) -> None:
    """Runs the plan generator (no compiler, no GPU) on an architecture + library and raises ValueError with the kernel family's
    limit that was hit (hidden width <= 64, <= 12 jet components, derivative order <= 6, d <= 4, tanh or Rational).  The Settings
    driver calls it before any data is loaded, so an unsupported configuration fails at start-up with an actionable message."""
    key = Hidden_Activation.strip().lower()
    if key in ("rat", "rational"):
        act = L.PDL_ACT_RATIONAL
    elif key == "tanh":
        act = L.PDL_ACT_TANH
    else:
        raise ValueError("Hidden Activation Function %r: the fused kernels implement Tanh and Rational (the reference's other "
                         "activations have no jet recurrences here, and there is no CPU fallback)" % Hidden_Activation)
    encs, terms = _library_key(Derivatives, LHS_Term, RHS_Terms)
    desc, keep = L.make_desc(int(d), [int(d)] + [int(w) for w in Hidden_Widths] + [1], act, L.PDL_MODE_COLLOCATION, encs,
                             [list(t) for t in terms])
    try:
        L.plan_source(desc)
    except L.NativeLibraryError as ex:
        raise ValueError("this architecture / library is outside the kernel family's limits (hidden width <= 64, <= 12 jet "
                         "components after closing the derivative set downward, derivative order <= 6, 2 <= d <= 4): %s" % ex) from None


def data_plan(U) -> Plan:
    widths, act = network_signature(U)
    key = (widths, act, L.PDL_MODE_DATA)
    if key not in _PLANS:
        _PLANS[key] = Plan(widths[0], widths, act, L.PDL_MODE_DATA, [(0, 0, 0, 0)], [[(0, 1)]])
    return _PLANS[key]


# ----------------------------------------------------------------------------------------------
# mask cache (the reference keeps Mask as a CPU bool tensor; the kernels want a device byte array)
# ----------------------------------------------------------------------------------------------
_MASKS: Dict[tuple, torch.Tensor] = {}


def device_mask(Mask, K: int, device: torch.device) -> torch.Tensor:
    if isinstance(Mask, torch.Tensor) and Mask.is_cuda:
        return Mask.to(torch.uint8).contiguous()
    m = np.asarray(Mask.detach().cpu().numpy() if isinstance(Mask, torch.Tensor) else Mask).astype(np.uint8).reshape(-1)
    assert m.size == K, "Mask must have one entry per RHS term"
    key = (m.tobytes(), device.index)
    if key not in _MASKS:
        _MASKS[key] = torch.from_numpy(m.copy()).to(device)
    return _MASKS[key]


# ----------------------------------------------------------------------------------------------
# autograd wrappers: loss and every gradient come out of ONE fused launch; backward only rescales
# ----------------------------------------------------------------------------------------------

def _split_flat(flat: torch.Tensor, shapes) -> List[torch.Tensor]:
    sizes = [math.prod(shp) for shp in shapes]
    return [piece.view(shp) for piece, shp in zip(flat.split(sizes), shapes)]


class _CollLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: Plan, points, mask_dev, inv_n, want_grad, xi, *params):
        dev = points.device
        n = points.shape[0]
        residual = torch.empty(n, dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float64, device=dev)
        gp = gx = None
        if want_grad:
            gp = torch.empty(plan.n_param_scalars, dtype=torch.float32, device=dev)
            gx = torch.empty(max(plan.K, 1), dtype=torch.float32, device=dev)
        plan.launch(points, params, xi, mask_dev, None, inv_n, want_grad, residual, None, gp, gx, stats)
        ctx.gp, ctx.gx, ctx.n_params, ctx.K = gp, gx, len(params), plan.K
        ctx.shapes = [p.shape for p in params]
        ctx.mark_non_differentiable(residual)
        return stats[0].to(torch.float32), residual

    @staticmethod
    def backward(ctx, g_loss, _g_res):
        if ctx.gp is None:
            raise RuntimeError("Coll_Loss was evaluated without gradients")
        grads = _split_flat(ctx.gp * g_loss, ctx.shapes)
        return (None, None, None, None, None, ctx.gx[:ctx.K] * g_loss, *grads)


class _DataLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: Plan, inputs, targets, inv_n, want_grad, *params):
        dev = inputs.device
        stats = torch.empty(4, dtype=torch.float64, device=dev)
        gp = torch.empty(plan.n_param_scalars, dtype=torch.float32, device=dev) if want_grad else None
        plan.launch(inputs, params, None, None, targets, inv_n, want_grad, None, None, gp, None, stats)
        ctx.gp = gp
        ctx.shapes = [p.shape for p in params]
        return stats[0].clone()                       # float64, like the reference (f32 prediction - f64 target)

    @staticmethod
    def backward(ctx, g_loss):
        if ctx.gp is None:
            raise RuntimeError("Data_Loss was evaluated without gradients")
        grads = _split_flat(ctx.gp * g_loss.to(torch.float32), ctx.shapes)
        return (None, None, None, None, None, *grads)


class _LpLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xi, mask_dev, p):
        out = torch.empty(1, dtype=torch.float32, device=xi.device)
        grad = torch.empty_like(xi)
        with torch.cuda.device(xi.device):
            L.check(L.lib().pdl_lp_loss(xi.data_ptr(), mask_dev.data_ptr(), int(xi.numel()), float(p), out.data_ptr(),
                                        grad.data_ptr(), torch.cuda.current_stream(xi.device).cuda_stream))
        ctx.grad = grad
        return out[0]

    @staticmethod
    def backward(ctx, g):
        return ctx.grad * g, None, None


class _L2LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *params):
        dev = params[0].device
        n = len(params)
        ptrs = (C.c_void_p * n)(*[p.data_ptr() for p in params])
        sizes = (C.c_int32 * n)(*[int(p.numel()) for p in params])
        total = sum(int(p.numel()) for p in params)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty(total, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().pdl_l2_squared_loss(ptrs, sizes, n, out.data_ptr(), grad.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream))
        ctx.grad = grad
        ctx.shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, g):
        return tuple(_split_flat(ctx.grad * g, ctx.shapes))


# ----------------------------------------------------------------------------------------------
# fused closure: the whole  sum_i [wD Data_i + wC Coll_i + wLp Lp + wL2 L2_i]  as ONE autograd node
# ----------------------------------------------------------------------------------------------

class ClosureSpec:
    """Everything that is not a differentiable tensor (kept out of autograd's argument handling)."""

    def __init__(self, U_List, Mask, Coll_Points_List, Inputs_List, Targets_List, Derivatives, LHS_Term, RHS_Terms, p, Weights,
                 inv_n_coll, inv_n_data, world, group=None):
        self.S = len(U_List)
        self.group = group
        self.coll_plans = [collocation_plan(U, Derivatives, LHS_Term, RHS_Terms) for U in U_List]
        self.data_plans = [data_plan(U) for U in U_List]
        self.n_tensors = [len(network_param_tensors(U)) for U in U_List]
        self.points = [_check_cuda_f32(P, "Coll_Points").detach() for P in Coll_Points_List]
        self.inputs = [_check_cuda_f32(X, "Inputs").detach() for X in Inputs_List]
        self.targets = []
        for X, Y in zip(self.inputs, Targets_List):
            if not Y.is_cuda:
                raise RuntimeError("Targets must be CUDA tensors: this engine has no CPU path")
            Y = Y.reshape(-1).to(torch.float64).contiguous()
            assert Y.numel() == X.shape[0]
            self.targets.append(Y)
        self.K = self.coll_plans[0].K
        self.mask_dev = device_mask(Mask, self.K, self.points[0].device)
        self.p, self.W, self.world = float(p), dict(Weights), int(world)
        self.inv_n_coll, self.inv_n_data = list(inv_n_coll), list(inv_n_data)


class _ClosureFn(torch.autograd.Function):
    """
    forward: launches Lp, and per network the fused collocation kernel, the fused data kernel and L2, then combines
    the flat gradients with the loss weights (one axpby3 pass per network).  Returns (total, scalars, *residuals):
    scalars = [coll_0..coll_{S-1}, data_*, l2_*, total_*, lp] (float64, device).

    All gradients of one closure live in ONE flat buffer  [network 0 | network 1 | ... | xi]  that the kernels write in place.
    backward rescales that buffer once, all-reduces it once when the closure runs data-parallel (spec.group: the single collective
    of the path) and hands autograd views of it: no split -> cat -> copy round trip, and the .grad tensors of all
    parameters share the reduced buffer.
    """

    @staticmethod
    def forward(ctx, spec: ClosureSpec, want_grad: bool, xi, *params):
        S, K, W, world = spec.S, spec.K, spec.W, spec.world
        dev = xi.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib = L.lib()
        lp = torch.empty(1, dtype=torch.float32, device=dev)
        g_lp = torch.empty(K, dtype=torch.float32, device=dev)
        stats_c = torch.empty((S, 4), dtype=torch.float64, device=dev)
        stats_d = torch.empty((S, 4), dtype=torch.float64, device=dev)
        l2 = torch.empty(S, dtype=torch.float32, device=dev)
        gx = torch.empty((S, max(K, 1)), dtype=torch.float32, device=dev) if want_grad else None
        n_scal = [pl.n_param_scalars for pl in spec.coll_plans]
        flat = torch.empty(sum(n_scal) + K, dtype=torch.float32, device=dev) if want_grad else None
        residuals, goff = [], 0
        with torch.cuda.device(dev):
            L.check(lib.pdl_lp_loss(xi.data_ptr(), spec.mask_dev.data_ptr(), K, spec.p, lp.data_ptr(), g_lp.data_ptr(), stream))
            off = 0
            for i in range(S):
                prm = params[off:off + spec.n_tensors[i]]
                off += spec.n_tensors[i]
                cp, dp = spec.coll_plans[i], spec.data_plans[i]
                P = cp.n_param_scalars
                res = torch.empty(spec.points[i].shape[0], dtype=torch.float32, device=dev)
                gc = flat[goff:goff + P] if want_grad else None
                goff += P
                gd = torch.empty(P, dtype=torch.float32, device=dev) if want_grad else None
                g2 = torch.empty(P, dtype=torch.float32, device=dev) if want_grad else None
                cp.launch(spec.points[i], prm, xi, spec.mask_dev, None, spec.inv_n_coll[i], want_grad, res, None, gc,
                          gx[i] if want_grad else None, stats_c[i])
                dp.launch(spec.inputs[i], prm, None, None, spec.targets[i], spec.inv_n_data[i], want_grad, None, None, gd, None,
                          stats_d[i])
                n = len(prm)
                ptrs = (C.c_void_p * n)(*[q.data_ptr() for q in prm])
                sizes = (C.c_int32 * n)(*[int(q.numel()) for q in prm])
                L.check(lib.pdl_l2_squared_loss(ptrs, sizes, n, l2[i:i + 1].data_ptr(), g2.data_ptr() if want_grad else None, stream))
                if want_grad:
                    L.check(lib.pdl_axpby3(gc.data_ptr(), P, gc.data_ptr(), W["Coll"], gd.data_ptr(), W["Data"], g2.data_ptr(),
                                           W["L2"] / world, stream))
                residuals.append(res)
        # scalar tail (totals, reported vector, summed total, weighted xi gradient): one launch
        scalars = torch.empty(4 * S + 1, dtype=torch.float64, device=dev)
        total = torch.empty(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.pdl_closure_tail(stats_c.data_ptr(), stats_d.data_ptr(), l2.data_ptr(), lp.data_ptr(),
                                         gx.data_ptr() if want_grad else None, g_lp.data_ptr(), S, K, max(K, 1),
                                         W["Data"], W["Coll"], W["Lp"] / world, W["L2"] / world, S * W["Lp"] / world, scalars.data_ptr(),
                                         total.data_ptr(), flat[goff:goff + K].data_ptr() if (want_grad and K > 0) else None, stream))
        ctx.flat, ctx.K, ctx.group = flat, K, spec.group
        ctx.shapes = [q.shape for q in params]
        for r in residuals:
            ctx.mark_non_differentiable(r)
        ctx.mark_non_differentiable(scalars)
        return (total[0], scalars, *residuals)

    @staticmethod
    def backward(ctx, g_total, _g_scalars, *_g_res):
        if ctx.flat is None:
            raise RuntimeError("the closure was evaluated without gradients")
        g = ctx.flat * g_total
        if ctx.group is not None:
            import torch.distributed as dist
            dist.all_reduce(g, group=ctx.group)
        n = g.numel() - ctx.K
        return (None, None, g[n:], *_split_flat(g[:n], ctx.shapes))


def closure_total(U_List, Xi, Mask, Coll_Points_List, Inputs_List, Targets_List, Derivatives, LHS_Term, RHS_Terms, p, Weights,
                  inv_n_coll=None, inv_n_data=None, world=1, group=None):
    """(total, scalars, residuals): see _ClosureFn.  `total` is differentiable w.r.t. Xi and every network parameter.
    With `group` (a torch.distributed process group of `world` ranks, each holding a shard of the rows) total.backward() leaves the
    all-reduced gradients in .grad."""
    assert torch.numel(Xi) == len(RHS_Terms)
    assert p > 0 and p < 2
    S = len(U_List)
    inv_c = inv_n_coll or [1.0 / max(int(P.shape[0]), 1) for P in Coll_Points_List]
    inv_d = inv_n_data or [1.0 / max(int(X.shape[0]), 1) for X in Inputs_List]
    spec = ClosureSpec(U_List, Mask, Coll_Points_List, Inputs_List, Targets_List, Derivatives, LHS_Term, RHS_Terms, p, Weights,
                       inv_c, inv_d, world, group)
    params = [_check_cuda_f32(q, "network parameter") for U in U_List for q in network_param_tensors(U)]
    xi = _check_cuda_f32(Xi, "Xi")
    want = torch.is_grad_enabled() and (xi.requires_grad or any(q.requires_grad for q in params))
    out = _ClosureFn.apply(spec, want, xi, *params)
    return out[0], out[1], list(out[2:2 + S])
