"""Pre-compile (no GPU needed) the collocation plans of the BASELINE configs with the current PDL_* knobs."""
import multiprocessing as mp
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g


def main():
    names = sys.argv[1:] or ["heat", "burgers", "kdv", "ks", "rd2d"]
    from pde_learn_b200 import configs, engine, _lib as L
    menu = []
    for name in names:
        d, _, derivs, lhs, rhs = configs.config(name)
        encs, terms = engine.library_key(derivs, lhs, rhs)
        for act in (L.PDL_ACT_TANH, L.PDL_ACT_RATIONAL):
            menu.append((d, [d] + [40] * 5 + [1], act, L.PDL_MODE_COLLOCATION, encs, [list(t) for t in terms]))
    with mp.get_context("spawn").Pool(min(8, len(menu))) as pool:
        pool.map(g._compile_one, menu)
    print("compiled", len(menu))


if __name__ == "__main__":
    main()
