"""The C ABI loads and exports every symbol include/pdelearn_b200.h declares; plan generation and its errors."""
import ctypes as C
import os
import re

import pytest

from pde_learn_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "pdelearn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pdl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 15
    lib = L.lib()
    for s in syms:
        assert hasattr(lib, s), "libpdelearn_b200.so does not export %s" % s
        assert s in L.PROTOTYPES, "no ctypes prototype for %s" % s
    assert lib.pdl_abi_version() == 3


def _heat_desc(widths=(2, 40, 40, 40, 40, 40, 1), act=L.PDL_ACT_TANH):
    encs = [(0, 0), (1, 0), (0, 1), (0, 2)]
    terms = [[(1, 1)], [(0, 1)], [(2, 1)], [(3, 1)], [(0, 2)]]
    return L.make_desc(2, list(widths), act, L.PDL_MODE_COLLOCATION, encs, terms)


def test_plan_source_contains_faa_di_bruno_terms():
    # KS set {U, U_t, U_x, U_xx, U_xxx, U_xxxx}: h_xxxx has the 5 multiset partitions of 4 with 1,6,3,4,1
    encs = [(0, 0), (1, 0), (0, 1), (0, 2), (0, 3), (0, 4)]
    terms = [[(1, 1)]] + [[(i, 1)] for i in (0, 2, 3, 4, 5)]
    desc, keep = L.make_desc(2, [2, 40, 40, 1], L.PDL_ACT_TANH, 0, encs, terms)
    src = L.plan_source(desc)
    line = [l for l in src.splitlines() if l.strip().startswith("h[5] =")][0]
    assert line.count("+") == 4
    for frag in ("s[4]*z[2]*z[2]*z[2]*z[2]", "6.f*s[3]*z[2]*z[2]*z[3]", "3.f*s[2]*z[3]*z[3]", "4.f*s[2]*z[2]*z[4]", "s[1]*z[5]"):
        assert frag in line, (frag, line)
    # tanh: sigma_3 = (1 - t^2)(6 t^2 - 2), sigma_5 needs the quartic 16 - 120 t^2 + 120 t^4
    assert "s[3] = c1 * (-2.f + t2 * 6.f);" in src
    assert "s[5] = c1 * (16.f + t2 * (-120.f + t2 * 120.f));" in src


def test_closure_adds_missing_lower_derivatives():
    # library uses only U and U_xx: the jet set must still contain U_x
    desc, keep = L.make_desc(2, [2, 8, 1], L.PDL_ACT_TANH, 0, [(0, 0), (0, 2)], [[(1, 1)], [(0, 1)]])
    src = L.plan_source(desc)
    assert "#define PDL_J 3" in src and "[1]=(0,1,0,0)" in src


@pytest.mark.parametrize("kwargs,msg", [
    (dict(d=5), "dimension"), (dict(widths=[2, 1]), "linear layers"), (dict(widths=[2, 80, 1]), "64"),
    (dict(act=7), "activation"), (dict(encs=[(0, 0), (0, 0, 1)]), "axis"), (dict(pow0=0), "powers")])
def test_plan_errors_are_loud(kwargs, msg):
    d = kwargs.get("d", 2)
    widths = kwargs.get("widths", [2, 16, 1])
    if "d" in kwargs:
        widths = [d] + widths[1:]
    encs = kwargs.get("encs", [(0, 0), (1, 0)])
    terms = [[(1, 1)], [(0, kwargs.get("pow0", 1))]]
    desc, keep = L.make_desc(d, widths, kwargs.get("act", 0), 0, encs, terms)
    need = C.c_size_t(0)
    rc = L.lib().pdl_plan_source(C.byref(desc), None, 0, C.byref(need))
    assert rc != 0
    assert msg in L.lib().pdl_last_error().decode()


def test_plan_create_compiles_for_sm100a_without_a_gpu():
    desc, keep = _heat_desc(widths=(2, 16, 16, 1))
    h = C.c_void_p()
    L.check(L.lib().pdl_plan_create(C.byref(desc), C.byref(h)))
    info = L.PlanInfo()
    L.check(L.lib().pdl_plan_info_get(h, C.byref(info)))
    assert info.n_jets == 4 and info.n_rhs_terms == 4 and info.d == 2
    assert info.n_param_scalars == 2 * 16 + 16 + 16 * 16 + 16 + 16 + 1
    assert info.flops_per_point == 6 * (16 * 16 * 4 + 16 * 4) + 4 * 2 * 16
    assert info.smem_bytes <= 227 * 1024 and info.threads_per_cta % 32 == 0
    assert L.lib().pdl_plan_workspace_bytes(h, 148) > 0
    # launching without the required buffers is refused before touching the device
    a = L.LaunchArgs()
    assert L.lib().pdl_loss_launch(h, C.byref(a)) != 0
    L.lib().pdl_plan_destroy(h)


def test_sampler_and_lp_argument_checks():
    import numpy as np
    b = np.array([1.0, 0.0], dtype=np.float64)            # lower > upper (Points.py:41-42)
    rc = L.lib().pdl_sample_points(b.ctypes.data_as(C.POINTER(C.c_double)), 1, 0, 0, 0, None, None)
    assert rc != 0 and "Bounds" in L.lib().pdl_last_error().decode()
    rc = L.lib().pdl_lp_loss(None, None, 3, 2.5, None, None, None)
    assert rc != 0 and "0 < p < 2" in L.lib().pdl_last_error().decode()


def test_build_menu_covers_the_golden_fixtures():
    """__graft_entry__.build() pre-compiles, for every golden fixture, the collocation plan and the Evaluate_Derivatives plan exactly
    as the host layer will request them (same derivative order), so the GPU box does not have to run nvcc for the parity tests."""
    import __graft_entry__ as G
    import pde_learn_b200 as P
    from pde_learn_b200 import engine as E
    from pde_learn_b200.Term import Term
    from tests.helpers import GOLDEN_CASES, GoldenCase
    import numpy as np

    def terms_of(g):
        from oracle import pde_oracle as O
        mk = lambda t: P.Term([P.Derivative(np.array(e)) for e, _ in t], [p for _, p in t])
        return [P.Derivative(np.array(e)) for e in O.distinct_encodings(g.lhs, g.rhs)], mk(g.lhs), [mk(t) for t in g.rhs]

    menu = {repr((m[0], list(m[1]), m[2], m[3], [tuple(e) for e in m[4]], [[tuple(x) for x in t] for t in m[5]])) for m in G._plan_menu()}
    for name in GOLDEN_CASES:
        g = GoldenCase(name)
        derivs, lhs, rhs = terms_of(g)
        act = L.PDL_ACT_RATIONAL if g.activation == "rational" else L.PDL_ACT_TANH
        for a, b in ((lhs, rhs), (Term([derivs[0]], [1]), [Term([D], [1]) for D in derivs[1:]])):
            encs, terms = E.library_key(derivs, a, b)
            key = repr((g.d, [int(v) for v in g.widths], act, L.PDL_MODE_COLLOCATION, [tuple(e) for e in encs], [[tuple(x) for x in t] for t in terms]))
            assert key in menu, (name, key)
