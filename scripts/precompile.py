"""Pre-compile (no GPU needed) the kernel plans of __graft_entry__._plan_menu() under the current PDL_* knobs, so that an A/B run
on the GPU box never waits for nvcc.  usage: PDL_CROSS_BF16=6 python scripts/precompile.py [5x40]   (5x40 = BASELINE workloads only)"""
import multiprocessing as mp
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

if __name__ == "__main__":
    menu = g._plan_menu()
    if len(sys.argv) > 1 and sys.argv[1] == "5x40":
        menu = [m for m in menu if m[1][1:-1] == [40] * 5 and m[3] == 0]
    with mp.get_context("spawn").Pool(min(8, os.cpu_count() or 1)) as pool:
        ok = pool.map(g._compile_one, menu)
    print("precompiled %d plans with %s" % (len(ok), {k: v for k, v in os.environ.items() if k.startswith("PDL_")}))
