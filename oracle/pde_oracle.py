"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A plain-PyTorch (CPU, float64 by default) restatement of the collocation-loss hot path of
punkduckable/PDE-LEARN.  It follows the *reference's own algorithm* -- nested `torch.autograd.grad`
through an ordinary MLP, one differentiation order at a time -- and NOT the jet/Taylor-mode algorithm the
CUDA kernels use, so the two are independent derivations of the same numbers.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import this file.  The product package (`pde_learn_b200/`) never does and has no CPU fallback.

Parity status: PINNED.  `tests/test_oracle_vs_golden.py` checks every function here against
  (a) golden vectors produced by importing the unmodified reference from /root/reference
      (`tests/golden/make_golden.py`, outputs committed as `tests/golden/*.npz`), and
  (b) the known answers of the reference's own unit tests (analytic polynomial derivatives,
      `Test/Test_Evaluate_Derivatives.py:33-195`; the hand-assembled collocation loss,
      `Test/Test_Loss.py:37-110`; the L^p closed forms, `Test/Test_Loss.py:180-223`).

Reference map (file:line under /root/reference):
  mlp_forward            <- Code/Classes/Network.py:287-312 (forward), :32-59 (Rational.forward)
  init_network           <- Code/Classes/Network.py:105-129 (Linear stack, Xavier-uniform W, zero b,
                            one Rational (a,b) per hidden layer, identity output activation)
  partial_derivative     <- Code/Evaluate_Derivatives.py:20-207 (t, then x, then y, then z; one
                            autograd.grad(create_graph=True) + column select per order)
  derivative_features    <- Code/Loss.py:138-181 (always from U itself: the child-reuse loop at
                            Loss.py:147 is `range(j, 0)` and never runs)
  term_value             <- Code/Loss.py:189-204 / :224-242 (product of powers)
  coll_loss              <- Code/Loss.py:66-248 (Mask[k]==True => term k contributes 0*Xi[k])
  data_loss              <- Code/Loss.py:23-62 (float32 prediction minus float64 target => float64 mean)
  lp_loss                <- Code/Loss.py:252-312 (IRLS surrogate, detached weights, delta = 1e-7)
  l2_squared_loss        <- Code/Loss.py:316-361 (W, b of every layer + Rational a, b)
  closure_losses         <- Code/Test_Train.py:131-190 (Lp added once per data set)
  targeted_cutoff_select <- Code/main.py:365-384 (|r| >= mean + 3*std over the random rows)
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

Encoding = Tuple[int, ...]                 # (n_t, n_x[, n_y[, n_z]])
SubTerm = Tuple[Encoding, int]             # (derivative encoding, power >= 1)
TermSpec = List[SubTerm]                   # product of sub-terms

RATIONAL_A_INIT = (0.0218, 0.5, 1.5957, 1.1915)      # Network.py:19-24
RATIONAL_B_INIT = (1.0, 0.0, 2.3830)                 # Network.py:25-29


# ----------------------------------------------------------------------------------------------
# Network
# ----------------------------------------------------------------------------------------------

@dataclass
class NetParams:
    """Flat description of one system-response MLP (what Network.py holds in nn.Modules)."""
    weights: List[torch.Tensor]                      # W_i [n_out, n_in]
    biases: List[torch.Tensor]                       # b_i [n_out]
    activation: str = "tanh"                         # "tanh" | "rational"
    rat_a: List[torch.Tensor] = field(default_factory=list)   # one [4] per hidden layer
    rat_b: List[torch.Tensor] = field(default_factory=list)   # one [3] per hidden layer

    def tensors(self) -> List[torch.Tensor]:
        """Same order as `list(Network.parameters())`: all Linear (W, b) pairs, then Rational (a, b)."""
        out: List[torch.Tensor] = []
        for W, b in zip(self.weights, self.biases):
            out += [W, b]
        for a, b in zip(self.rat_a, self.rat_b):
            out += [a, b]
        return out

    def to(self, dtype: torch.dtype, requires_grad: bool = False) -> "NetParams":
        conv = lambda t: t.detach().clone().to(dtype).requires_grad_(requires_grad)
        return NetParams([conv(w) for w in self.weights], [conv(b) for b in self.biases], self.activation,
                         [conv(a) for a in self.rat_a], [conv(b) for b in self.rat_b])


def init_network(widths: Sequence[int], activation: str, seed: int,
                 dtype: torch.dtype = torch.float32) -> NetParams:
    """Xavier-uniform weights, zero biases, Rational coefficients at their defaults (Network.py:117-129)."""
    gen = torch.Generator().manual_seed(seed)
    Ws, bs, ra, rb = [], [], [], []
    for n_in, n_out in zip(widths[:-1], widths[1:]):
        bound = math.sqrt(6.0 / (n_in + n_out))
        Ws.append(((torch.rand(n_out, n_in, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype))
        bs.append(torch.zeros(n_out, dtype=dtype))
    act = activation.strip().lower()
    if act == "rational":
        for _ in range(len(widths) - 2):
            ra.append(torch.tensor(RATIONAL_A_INIT, dtype=dtype))
            rb.append(torch.tensor(RATIONAL_B_INIT, dtype=dtype))
    return NetParams(Ws, bs, act, ra, rb)


def mlp_forward(net: NetParams, X: torch.Tensor) -> torch.Tensor:
    """act(Linear) for every hidden layer, plain Linear on the output layer.  Returns [B, n_out]."""
    H = len(net.weights) - 1
    for i in range(H + 1):
        X = X @ net.weights[i].t() + net.biases[i]
        if i < H:
            if net.activation == "tanh":
                X = torch.tanh(X)
            elif net.activation == "rational":
                a, b = net.rat_a[i], net.rat_b[i]
                num = a[0] + X * (a[1] + X * (a[2] + a[3] * X))
                den = b[0] + X * (b[1] + b[2] * X)
                X = num / den
            else:
                raise ValueError("oracle supports tanh and rational, got %r" % net.activation)
    return X


# ----------------------------------------------------------------------------------------------
# Derivatives by nested reverse-mode differentiation (the reference's algorithm)
# ----------------------------------------------------------------------------------------------

def partial_derivative(values: torch.Tensor, coords: torch.Tensor, extra: Sequence[int]) -> torch.Tensor:
    """
    Apply d_t^{extra[0]} d_x^{extra[1]} ... to `values` (shape [B], graph-attached to `coords` [B, d]).
    Axis order t, x, y, z and one gradient call per order, as Evaluate_Derivatives.py:104-205.
    """
    cur = values
    for axis, order in enumerate(extra):
        for _ in range(int(order)):
            g = torch.autograd.grad(cur, coords, grad_outputs=torch.ones_like(cur),
                                    retain_graph=True, create_graph=True)[0]
            cur = g[:, axis].reshape(-1)
    return cur


def derivative_features(u_fn: Callable[[torch.Tensor], torch.Tensor], coords: torch.Tensor,
                        encodings: Sequence[Encoding]) -> Dict[Encoding, torch.Tensor]:
    """{encoding: D_enc U at every row of coords}; every derivative is taken from U itself (Loss.py:166-181)."""
    coords.requires_grad_(True)
    u = u_fn(coords).reshape(-1)
    return {tuple(int(e) for e in enc): partial_derivative(u, coords, enc) for enc in encodings}


def term_value(term: TermSpec, feats: Dict[Encoding, torch.Tensor], like: torch.Tensor) -> torch.Tensor:
    out = torch.ones_like(like)
    for enc, power in term:
        out = out * torch.pow(feats[tuple(enc)], power)
    return out


def distinct_encodings(lhs: TermSpec, rhs: Sequence[TermSpec]) -> List[Encoding]:
    """Distinct encodings, LHS first, stably sorted by total order (Library_Reader.py:242-277)."""
    seen: List[Encoding] = []
    for term in [lhs] + list(rhs):
        for enc, _ in term:
            e = tuple(int(v) for v in enc)
            if e not in seen:
                seen.append(e)
    return sorted(seen, key=lambda e: sum(e))


# ----------------------------------------------------------------------------------------------
# Losses
# ----------------------------------------------------------------------------------------------

def coll_loss(u_fn: Callable[[torch.Tensor], torch.Tensor], xi: torch.Tensor, mask: Sequence[bool],
              coords: torch.Tensor, lhs: TermSpec, rhs: Sequence[TermSpec],
              encodings: Optional[Sequence[Encoding]] = None
              ) -> Tuple[torch.Tensor, torch.Tensor, Dict[Encoding, torch.Tensor]]:
    """mean((T_0(U) - sum_k xi_k T_k(U))^2), the residual vector, and the derivative features."""
    assert xi.numel() == len(rhs)                                   # Loss.py:124
    if encodings is None:
        encodings = distinct_encodings(lhs, rhs)
    feats = derivative_features(u_fn, coords, encodings)
    like = next(iter(feats.values()))
    b_u = term_value(lhs, feats, like)
    l_xi = torch.zeros_like(b_u)
    for k, term in enumerate(rhs):
        if bool(mask[k]):
            l_xi = l_xi + 0.0 * xi[k]                               # keeps xi[k] in the graph, grad exactly 0
        else:
            l_xi = l_xi + term_value(term, feats, like) * xi[k]
    residual = b_u - l_xi
    return (residual ** 2).mean(), residual, feats


def data_loss(u_fn: Callable[[torch.Tensor], torch.Tensor], inputs: torch.Tensor,
              targets: torch.Tensor) -> torch.Tensor:
    """Mean square error; dtype follows torch promotion (f32 prediction - f64 target => f64)."""
    return ((u_fn(inputs).squeeze() - targets) ** 2).mean()


def lp_loss(xi: torch.Tensor, mask: Sequence[bool], p: float) -> torch.Tensor:
    """sum_k w_k xi_k^2 with w_k = 1/max(1e-7, |xi_k|^(2-p)) detached, w_k = 0 where masked."""
    assert 0.0 < p < 2.0                                            # Loss.py:279
    delta = 1e-7
    w = torch.zeros_like(xi.detach())
    for k in range(xi.numel()):
        if not bool(mask[k]):
            w[k] = 1.0 / max(delta, abs(float(xi[k].detach())) ** (2.0 - p))
    return (w * xi * xi).sum()


def l2_squared_loss(net: NetParams) -> torch.Tensor:
    """Sum of squares of every trainable scalar of the network; shape [1] like the reference."""
    acc = torch.zeros(1, dtype=net.weights[0].dtype)
    for t in net.tensors():
        acc = acc + (t * t).sum()
    return acc


@dataclass
class ClosureResult:
    total: torch.Tensor
    coll: List[float]
    data: List[float]
    l2: List[float]
    lp: float
    totals: List[float]
    residuals: List[torch.Tensor]
    grads_nets: List[List[torch.Tensor]]
    grad_xi: torch.Tensor


def closure_losses(nets: Sequence[NetParams], xi: torch.Tensor, mask: Sequence[bool],
                   coll_points: Sequence[torch.Tensor], inputs: Sequence[torch.Tensor],
                   targets: Sequence[torch.Tensor], lhs: TermSpec, rhs: Sequence[TermSpec],
                   p: float, weights: Dict[str, float], backward: bool = True) -> ClosureResult:
    """
    One Closure() of Test_Train.py:131-190: per data set  wD*Data + wC*Coll + wLp*Lp + wL2*L2, summed
    over data sets (so Lp enters once per data set), then backward.  `nets`/`xi` must require grad.
    """
    lp = lp_loss(xi, mask, p)
    total = None
    coll_l, data_l, l2_l, tot_l, res_l = [], [], [], [], []
    for net, pts, x_in, y in zip(nets, coll_points, inputs, targets):
        u_fn = lambda X, net=net: mlp_forward(net, X)
        c, r, _ = coll_loss(u_fn, xi, mask, pts, lhs, rhs)
        d = data_loss(u_fn, x_in, y)
        l2 = l2_squared_loss(net)
        t = weights["Data"] * d + weights["Coll"] * c + weights["Lp"] * lp + weights["L2"] * l2
        coll_l.append(float(c)); data_l.append(float(d)); l2_l.append(float(l2)); tot_l.append(float(t))
        res_l.append(r.detach())
        total = t if total is None else total + t
    grads_nets: List[List[torch.Tensor]] = []
    grad_xi = torch.zeros_like(xi)
    if backward:
        leaves = [t for net in nets for t in net.tensors()] + [xi]
        grads = torch.autograd.grad(total.sum(), leaves, allow_unused=True)
        it = iter(grads)
        for net in nets:
            grads_nets.append([g if g is not None else torch.zeros_like(t)
                               for t, g in zip(net.tensors(), it)])
        g = next(it)
        grad_xi = g if g is not None else grad_xi
    return ClosureResult(total.detach().reshape(-1), coll_l, data_l, l2_l, float(lp), tot_l, res_l,
                         grads_nets, grad_xi)


# ----------------------------------------------------------------------------------------------
# Targeted collocation points
# ----------------------------------------------------------------------------------------------

def targeted_cutoff_select(points: torch.Tensor, residual: torch.Tensor, num_random: int
                           ) -> Tuple[float, torch.Tensor]:
    """cutoff = mean + 3*std (unbiased) of |r| over the first num_random rows; keep rows with |r| >= cutoff."""
    a = residual.detach().abs()
    head = a[:num_random]
    cutoff = float(head.mean() + 3.0 * head.std())
    return cutoff, points[a >= cutoff]


# ----------------------------------------------------------------------------------------------
# Libraries used by tests / bench
# ----------------------------------------------------------------------------------------------

def monomials(symbols: Sequence[Encoding], degree: int) -> List[TermSpec]:
    """All monomials of exactly `degree` in the given derivative symbols, as TermSpecs."""
    from itertools import combinations_with_replacement
    out = []
    for combo in combinations_with_replacement(range(len(symbols)), degree):
        term: TermSpec = []
        for idx in sorted(set(combo)):
            term.append((symbols[idx], combo.count(idx)))
        out.append(term)
    return out


def config_library(name: str) -> Tuple[int, TermSpec, List[TermSpec], List[Tuple[float, float]]]:
    """(d, LHS, RHS terms, bounds) for the BASELINE.json configs, as defined in SURVEY.md section 8d."""
    U, Ut, Ux, Uxx, Uxxx, Uxxxx = (0, 0), (1, 0), (0, 1), (0, 2), (0, 3), (0, 4)
    if name == "burgers":          # shipped Library.txt:48-70
        rhs = [[(U, 1)], [(Ux, 1)], [(Uxx, 1)], [(Uxxx, 1)],
               [(U, 2)], [(Ux, 1), (U, 1)], [(Uxx, 1), (U, 1)], [(Ux, 2)],
               [(U, 3)], [(Ux, 1), (U, 2)], [(Uxx, 1), (U, 2)], [(Ux, 2), (U, 1)],
               [(U, 4)], [(Ux, 1), (U, 3)], [(Uxx, 1), (U, 3)], [(Ux, 2), (U, 2)], [(Ux, 3), (U, 1)]]
        return 2, [(Ut, 1)], rhs, [(0.0, 10.0), (-8.0, 8.0)]
    if name == "heat":
        s = [U, Ux, Uxx]
        return 2, [(Ut, 1)], monomials(s, 1) + monomials(s, 2), [(0.0, 10.0), (0.0, 10.0)]
    if name == "kdv":
        s = [U, Ux, Uxx, Uxxx]
        return 2, [(Ut, 1)], monomials(s, 1) + monomials(s, 2), [(0.0, 40.0), (-20.0, 20.0)]
    if name == "ks":
        s = [U, Ux, Uxx, Uxxx, Uxxxx]
        return (2, [(Ut, 1)], monomials(s, 1) + monomials(s, 2) + monomials([U, Ux, Uxx], 3),
                [(0.0, 50.0), (-10.0, 10.0)])
    if name == "rd2d":
        V, Vt, Vx, Vy, Vxx, Vyy = (0, 0), (1, 0), (0, 1), (0, 0, 1), (0, 2), (0, 0, 2)
        s = [V, Vx, Vy, Vxx, Vyy]
        return (3, [(Vt, 1)], monomials(s, 1) + monomials(s, 2) + [[(V, 3)]],
                [(0.0, 5.0), (-5.0, 5.0), (-5.0, 5.0)])
    raise KeyError(name)


# ----------------------------------------------------------------------------------------------
# Point sampler: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123
# library's philox4x32 with 10 rounds).  The reference samples with Python's random.uniform (Code/Points.py:45-57),
# which a counter-based device sampler cannot reproduce; what is pinned here is the published block function, by
# Random123's own known-answer vectors (tests/test_oracle_vs_golden.py), and the word -> coordinate map the CUDA
# sampler documents in include/pdelearn_b200.h.
# ----------------------------------------------------------------------------------------------
PHILOX_KAT = [   # (counter[4], key[2]) -> output[4]   (Random123 tests/kat_vectors, philox4x32 10 rounds)
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(counter, key):
    """counter: uint32 [n, 4], key: uint32 [n, 2] -> uint32 [n, 4] (numpy, vectorised over n)."""
    import numpy as np
    c = np.asarray(counter, dtype=np.uint64).reshape(-1, 4).copy()
    k = np.asarray(key, dtype=np.uint64).reshape(-1, 2).copy()
    mask = np.uint64(0xFFFFFFFF)
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    s32 = np.uint64(32)
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> s32, p0 & mask, p1 >> s32, p1 & mask
        c = np.stack([hi1 ^ c[:, 1] ^ k[:, 0], lo1, hi0 ^ c[:, 3] ^ k[:, 1], lo0], axis=1)
        k = np.stack([(k[:, 0] + W0) & mask, (k[:, 1] + W1) & mask], axis=1)
    return c.astype(np.uint32)


def sample_points(bounds, n: int, seed: int, first_index: int = 0):
    """The CUDA sampler restated: row i uses counter = {(first_index + i) low, high, 0, 0}, key = {seed low, high}; word j becomes
    coordinate j = min(fma(float(word >> 8) * 2^-24, hi_j - lo_j, lo_j), hi_j) in float32.  Returns float32 [n, d]."""
    import numpy as np
    b = np.asarray(bounds, dtype=np.float64).reshape(-1, 2)
    d = b.shape[0]
    idx = np.arange(n, dtype=np.uint64) + np.uint64(first_index)
    ctr = np.zeros((n, 4), dtype=np.uint64)
    ctr[:, 0], ctr[:, 1] = idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32)
    key = np.empty((n, 2), dtype=np.uint64)
    key[:, 0], key[:, 1] = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    w = philox4x32_10(ctr, key)
    lo32, hi32 = b[:, 0].astype(np.float32), b[:, 1].astype(np.float32)
    span = (hi32 - lo32).astype(np.float32)
    u = (w[:, :d] >> np.uint32(8)).astype(np.float64) * 2.0 ** -24            # exact
    # fused multiply-add in float32 = the exactly computed product plus lo, rounded once; float64 holds the 48-bit product exactly
    out = (u * span.astype(np.float64) + lo32.astype(np.float64)).astype(np.float32)
    return np.minimum(out, hi32[None, :])
