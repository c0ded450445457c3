#!/usr/bin/env python
"""
The reference arm of bench.py: times the UNMODIFIED reference (baseline/_ref/Code, installed by baseline/install_ref.py; falls
back to /root/reference/Code when that exists) on the host cores through its own public entry point and stock code path,

    Code/Test_Train.py:22   Training(U_List, Xi, Mask, Coll_Points_List, Inputs_List, Targets_List, Derivatives, LHS_Term,
                                     RHS_Terms, p, Weights, Optimizer, Device)

i.e. one closure = Lp_Loss + per network {Coll_Loss, Data_Loss, L2_Squared_Loss} -> weighted total -> backward
(Code/Test_Train.py:131-190) plus the Adam update Training always performs (:193; microseconds next to the closure).
Nothing from pde_learn_b200/ or oracle/ is imported here.  The workload definitions below restate SURVEY.md section 8(d) as plain
data (tests/test_host_logic.py checks them against pde_learn_b200.configs).

    python baseline/reference_arm.py --workload heat --activation tanh --rows 20000 --steps 5 --warmup 2     # prints one JSON object
"""
import argparse
import json
import os
import sys
import time
from itertools import combinations_with_replacement

HERE = os.path.dirname(os.path.abspath(__file__))
N_DATA = 10_000
WEIGHTS = {"Data": 1.0, "Coll": 1.0, "Lp": 2e-4, "L2": 5e-6}
P_EXP = 0.1


def find_reference():
    for base in (os.path.join(HERE, "_ref"), os.environ.get("PDL_REFERENCE_DIR", "/root/reference")):
        if base and os.path.isfile(os.path.join(base, "Code", "Loss.py")):
            return base
    return None


def _monomials(symbols, degree):
    out = []
    for combo in combinations_with_replacement(range(len(symbols)), degree):
        idx = sorted(set(combo))
        out.append([(symbols[i], combo.count(i)) for i in idx])
    return out


def workload(name):
    """(d, bounds, lhs, rhs, networks): a term is a list of (encoding, power) -- SURVEY.md 8(d) configs 1-5."""
    U, Ut, Ux, Uxx, Uxxx, Uxxxx = (0, 0), (1, 0), (0, 1), (0, 2), (0, 3), (0, 4)
    lhs = [(Ut, 1)]
    if name == "burgers":        # the shipped Library.txt:48-70
        rhs = [[(U, 1)], [(Ux, 1)], [(Uxx, 1)], [(Uxxx, 1)], [(U, 2)], [(Ux, 1), (U, 1)], [(Uxx, 1), (U, 1)], [(Ux, 2)], [(U, 3)],
               [(Ux, 1), (U, 2)], [(Uxx, 1), (U, 2)], [(Ux, 2), (U, 1)], [(U, 4)], [(Ux, 1), (U, 3)], [(Uxx, 1), (U, 3)],
               [(Ux, 2), (U, 2)], [(Ux, 3), (U, 1)]]
        return 2, [(0, 10), (-8, 8)], lhs, rhs, 1
    if name == "heat":
        s = [U, Ux, Uxx]
        return 2, [(0, 10), (0, 10)], lhs, _monomials(s, 1) + _monomials(s, 2), 1
    if name == "kdv":
        s = [U, Ux, Uxx, Uxxx]
        return 2, [(0, 40), (-20, 20)], lhs, _monomials(s, 1) + _monomials(s, 2), 1
    if name == "ks":
        s = [U, Ux, Uxx, Uxxx, Uxxxx]
        return 2, [(0, 50), (-10, 10)], lhs, _monomials(s, 1) + _monomials(s, 2) + _monomials([U, Ux, Uxx], 3), 1
    if name == "rd2d":
        V, Vx, Vy, Vxx, Vyy = (0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 2, 0), (0, 0, 2)
        s = [V, Vx, Vy, Vxx, Vyy]
        return 3, [(0, 5), (-5, 5), (-5, 5)], [((1, 0, 0), 1)], _monomials(s, 1) + _monomials(s, 2) + [[(V, 3)]], 2
    raise KeyError(name)


def run(workload_name, activation, rows, steps, warmup, max_seconds=150.0):
    base = find_reference()
    if base is None:
        return {"unavailable": "baseline/_ref/Code missing and /root/reference absent (run baseline/install_ref.py in the build container)"}
    code = os.path.join(base, "Code")
    for p in (os.path.join(code, "Readers"), os.path.join(code, "Classes"), code):
        sys.path.insert(0, p)
    import numpy
    import torch
    from Derivative import Derivative                    # reference classes (Code/Classes)
    from Term import Term
    from Network import Network
    from Test_Train import Training                      # reference entry point (Code/Test_Train.py:22)

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d, bounds, lhs, rhs, S = workload(workload_name)
    K = len(rhs)
    encs = []
    for t in [lhs] + rhs:
        for e, _ in t:
            if e not in encs:
                encs.append(e)
    encs.sort(key=lambda e: sum(e))
    mk = lambda t: Term(Derivatives=[Derivative(Encoding=numpy.array(e)) for e, _ in t], Powers=[int(p) for _, p in t])
    Derivatives = [Derivative(Encoding=numpy.array(e)) for e in encs]
    LHS, RHS = mk(lhs), [mk(t) for t in rhs]
    act = "Tanh" if activation == "tanh" else "Rational"
    U_List = []
    for i in range(S):
        torch.manual_seed(i)
        U_List.append(Network(Widths=[d] + [40] * 5 + [1], Hidden_Activation=act, Output_Activation="None"))
    gen = torch.Generator().manual_seed(0)
    Xi = (0.1 * torch.randn(K, generator=gen)).requires_grad_(True)
    Mask = torch.zeros(K, dtype=torch.bool)
    lo = torch.tensor([b[0] for b in bounds], dtype=torch.float32)
    hi = torch.tensor([b[1] for b in bounds], dtype=torch.float32)
    pts = [(lo + (hi - lo) * torch.rand(rows, d, generator=gen)) for _ in range(S)]
    x_in = [(lo + (hi - lo) * torch.rand(N_DATA, d, generator=gen)) for _ in range(S)]
    y = [torch.randn(N_DATA, generator=gen, dtype=torch.float64) for _ in range(S)]
    params = [q for U in U_List for q in U.parameters()] + [Xi]
    opt = torch.optim.Adam(params, lr=1e-3)
    times, t_start = [], time.perf_counter()
    done_warm = 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = Training(U_List=U_List, Xi=Xi, Mask=Mask, Coll_Points_List=[p.clone() for p in pts], Inputs_List=x_in, Targets_List=y,
                       Derivatives=Derivatives, LHS_Term=LHS, RHS_Terms=RHS, p=P_EXP, Weights=WEIGHTS, Optimizer=opt,
                       Device=torch.device("cpu"))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        else:
            done_warm += 1
        # bounded sample: stop early (after at least one timed step) if the run would exceed the time budget
        if times and (time.perf_counter() - t_start) + dt > max_seconds:
            break
    assert all(numpy.isfinite(v) for v in out["Total Losses"])
    total_rows = S * rows * len(times)
    return {"value": total_rows / sum(times), "best": S * rows / min(times), "unit": "points/s", "cores": cores, "kind": "reference",
            "steps": len(times), "warmup": done_warm, "ms_per_step": 1e3 * sum(times) / len(times), "rows": rows, "networks": S,
            "reference_dir": os.path.relpath(base, os.path.dirname(HERE)) if base.startswith(os.path.dirname(HERE)) else base,
            "sample": "%d call(s) of the unmodified reference Training() (Code/Test_Train.py:22; closure + Adam step) on %d rows x %d "
                      "network(s) (+%d data rows), float32, torch %s, %d threads"
                      % (len(times), rows, S, N_DATA, torch.__version__, cores)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="heat")
    ap.add_argument("--activation", default="tanh")
    ap.add_argument("--rows", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--max-seconds", type=float, default=150.0)
    a = ap.parse_args()
    print(json.dumps(run(a.workload, a.activation, a.rows, a.steps, a.warmup, a.max_seconds)), flush=True)
