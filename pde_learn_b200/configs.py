"""
The BASELINE.json workloads as (d, bounds, Derivatives, LHS_Term, RHS_Terms), following SURVEY.md section 8d.
Product-side definitions used by bench.py, __graft_entry__ and the tests (the oracle keeps its own copy).
"""
from itertools import combinations_with_replacement
from typing import List, Tuple

import numpy

from .Derivative import Derivative
from .Term import Term

SHIPPED_LIBRARY = """
D_t U
U
D_x U
D_x^2 U
D_x^3 U
(U)^2
(D_x U)*U
(D_x^2 U)*U
(D_x U)^2
(U)^3
(D_x U)*(U)^2
(D_x^2 U)*U^2
(D_x U)^2*U
(U)^4
(D_x U)*(U)^3
(D_x^2 U)*(U)^3
(D_x U)^2*(U)^2
(D_x U)^3*U
"""     # the reference's shipped Library.txt:48-70 (Burgers bring-up library)


def _D(*enc):
    return Derivative(Encoding=numpy.array(enc, dtype=numpy.int32))


def _monomials(symbols: List[Derivative], degree: int) -> List[Term]:
    out = []
    for combo in combinations_with_replacement(range(len(symbols)), degree):
        idx = sorted(set(combo))
        out.append(Term(Derivatives=[symbols[i] for i in idx], Powers=[combo.count(i) for i in idx]))
    return out


def _finish(d, bounds, lhs, rhs):
    seen, derivs = set(), []
    for T in [lhs] + rhs:
        for D in T.Derivatives:
            k = tuple(D.Encoding.tolist())
            if k not in seen:
                seen.add(k); derivs.append(D)
    derivs.sort(key=lambda D: D.Order)
    return d, numpy.array(bounds, dtype=numpy.float64), derivs, lhs, rhs


def config(name: str) -> Tuple[int, numpy.ndarray, List[Derivative], Term, List[Term]]:
    U, Ut, Ux, Uxx, Uxxx, Uxxxx = _D(0, 0), _D(1, 0), _D(0, 1), _D(0, 2), _D(0, 3), _D(0, 4)
    lhs = Term([Ut], [1])
    if name == "burgers":
        from .Library_Reader import Parse_Term
        terms = [Parse_Term(l) for l in SHIPPED_LIBRARY.strip().splitlines()]
        return _finish(2, [(0, 10), (-8, 8)], terms[0], terms[1:])
    if name == "heat":
        s = [U, Ux, Uxx]
        return _finish(2, [(0, 10), (0, 10)], lhs, _monomials(s, 1) + _monomials(s, 2))
    if name == "kdv":
        s = [U, Ux, Uxx, Uxxx]
        return _finish(2, [(0, 40), (-20, 20)], lhs, _monomials(s, 1) + _monomials(s, 2))
    if name == "ks":
        s = [U, Ux, Uxx, Uxxx, Uxxxx]
        return _finish(2, [(0, 50), (-10, 10)], lhs, _monomials(s, 1) + _monomials(s, 2) + _monomials([U, Ux, Uxx], 3))
    if name == "rd2d":
        V, Vx, Vy, Vxx, Vyy = _D(0, 0), _D(0, 1), _D(0, 0, 1), _D(0, 2), _D(0, 0, 2)
        s = [V, Vx, Vy, Vxx, Vyy]
        return _finish(3, [(0, 5), (-5, 5), (-5, 5)], lhs, _monomials(s, 1) + _monomials(s, 2) + [Term([V], [3])])
    raise KeyError(name)


CONFIG_POINTS = {"burgers": 10_000, "heat": 1 << 20, "kdv": 1 << 22, "ks": 1 << 24, "rd2d": 1 << 26}
CONFIG_NETWORKS = {"burgers": 1, "heat": 1, "kdv": 1, "ks": 1, "rd2d": 2}
