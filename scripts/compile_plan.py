"""Compile (no GPU needed) the plan of one workload with the current PDL_* knobs and print the ptxas resource lines.
usage: python scripts/compile_plan.py heat|burgers|kdv|ks|rd2d [Tanh|Rational]"""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_learn_b200 import _lib as L
from pde_learn_b200 import configs

w = sys.argv[1] if len(sys.argv) > 1 else "heat"
act = sys.argv[2] if len(sys.argv) > 2 else "Tanh"
from pde_learn_b200 import engine as E
d, bounds, derivs, lhs, rhs = configs.config(w)
encs, terms = E.library_key(derivs, lhs, rhs)
desc, keep = L.make_desc(d, [d, 40, 40, 40, 40, 40, 1], 0 if act == "Tanh" else 1, 0, encs, terms)
src = L.plan_source(desc)
csrc = os.path.join(os.path.dirname(os.path.abspath(L.__file__)), "csrc")
with tempfile.TemporaryDirectory() as td:
    cu = os.path.join(td, "plan.cu")
    open(cu, "w").write(src)
    r = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xptxas", "-v", "--shared",
                        "-Xcompiler", "-fPIC", "-I" + csrc] + os.environ.get("PDL_NVCC_FLAGS", "").split() + ["-o", os.path.join(td, "p.so"), cu], capture_output=True, text=True)
    lines = (r.stdout + r.stderr).splitlines()
    for i, l in enumerate(lines):
        if "pdl_main_kernel" in l and "Compiling" in l or "registers" in l or "spill" in l:
            print(l.strip()[:200])
    if len(sys.argv) > 3:
        import shutil
        shutil.copy(os.path.join(td, "p.so"), sys.argv[3])
    if r.returncode:
        print("\n".join(lines[-20:]))
