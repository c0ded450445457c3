"""
Read_Library(File_Path) -> (Derivatives, LHS_Term, RHS_Terms): parser for the reference's Library.txt
format (Code/Readers/Library_Reader.py:201-277; grammar documented in Library.txt:1-45).

  * '#' starts a comment, blank lines are skipped; the first term is the LHS, every further line one RHS term;
  * a term is a '*'-separated product of sub-terms;  a sub-term is  [(]D_s[^k] D_s[^k] ... U[)][^p]
    with s in {t,x,y,z}; no derivative token means the identity, encoded [0,0] whatever the dimension;
  * encodings are trimmed of trailing zero y/z slots (so D_x U -> [0,1] even in a 2-D problem);
  * the distinct derivatives are returned LHS-first, stably sorted by total order.
"""
from typing import List, Tuple

import numpy

from .Derivative import Derivative
from .Term import Term

_AXIS = {"t": 0, "x": 1, "y": 2, "z": 3}


def _parse_sub_term(text: str) -> Tuple[Derivative, int]:
    s = text.strip()
    pos = max(s.rfind("U"), s.rfind("u"))
    if pos < 0:
        raise ValueError("sub-term without U: %r" % text)
    head, tail = s[:pos], s[pos + 1:]
    power = 1
    if "^" in tail:
        power = int(tail.split("^", 1)[1].strip().strip(")").strip())
    enc = [0, 0, 0, 0]
    for tok in head.replace("(", " ").split():
        if not tok.startswith("D_"):
            raise ValueError("bad derivative token %r in %r" % (tok, text))
        body = tok[2:]
        order = 1
        if "^" in body:
            body, o = body.split("^", 1)
            order = int(o)
        if body.lower() not in _AXIS:
            raise ValueError("unknown differentiation variable %r in %r" % (body, text))
        enc[_AXIS[body.lower()]] += order
    n = 4
    while n > 2 and enc[n - 1] == 0:
        n -= 1
    return Derivative(Encoding=numpy.array(enc[:n], dtype=numpy.int32)), power


def Parse_Term(line: str) -> Term:
    derivs, powers = [], []
    for piece in line.split("*"):
        D, p = _parse_sub_term(piece)
        derivs.append(D); powers.append(p)
    return Term(Derivatives=derivs, Powers=powers)


def Read_Library(File_Path: str) -> Tuple[List[Derivative], Term, List[Term]]:
    terms: List[Term] = []
    with open(File_Path, "r") as f:
        for raw in f:
            line = raw.split("#", 1)[0].strip()
            if line:
                terms.append(Parse_Term(line))
    if not terms:
        raise ValueError("no terms in %s" % File_Path)
    seen, derivs = set(), []
    for T in terms:
        for D in T.Derivatives:
            key = tuple(D.Encoding.tolist())
            if key not in seen:
                seen.add(key); derivs.append(D)
    derivs.sort(key=lambda D: D.Order)          # stable, like the reference
    return derivs, terms[0], terms[1:]
