"""
TEST INFRASTRUCTURE ONLY: compiles the generated plan source as plain C++ (-DPDL_EMULATE) and runs the
kernel lane by lane on the CPU.  Used by the `-m "not gpu"` tests to validate the exact kernel source
(indexing, Faa-di-Bruno code generation, adjoint) against the oracle without a GPU.  The product never
loads these modules.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

from pde_learn_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pde_learn_b200", "csrc")
EMU_DIR = os.path.join(ROOT, "tests", "_emu_build")
EXTRA_FLAGS = []          # e.g. sanitizers (tests/test_emulation_asan.py)


class KernelParams(C.Structure):
    _fields_ = [("points", C.c_void_p), ("n_points", C.c_longlong), ("prm", C.c_void_p * 40), ("xi", C.c_void_p),
                ("mask", C.c_void_p), ("targets", C.c_void_p), ("inv_n", C.c_double), ("residual", C.c_void_p),
                ("features", C.c_void_p), ("zscratch", C.c_void_p), ("wpriv", C.c_void_p), ("cta_out", C.c_void_p),
                ("cta_stats", C.c_void_p), ("n_points_dev", C.c_void_p)]


def build_emu(desc):
    src = L.plan_source(desc)
    hdr = (open(os.path.join(CSRC, "jet_mlp_kernel.cuh")).read() + open(os.path.join(CSRC, "jet_mlp_kernel_mma.cuh")).read() +
           open(os.path.join(CSRC, "jet_common.cuh")).read())
    tag = hashlib.sha1((src + hdr).encode()).hexdigest()[:16]
    os.makedirs(EMU_DIR, exist_ok=True)
    so = os.path.join(EMU_DIR, "emu_%s.so" % tag)
    if not os.path.exists(so):
        cpp = os.path.join(EMU_DIR, "emu_%s.cpp" % tag)
        with open(cpp, "w") as f:
            f.write(src)
        cmd = ["g++", "-O1", "-std=c++17", "-DPDL_EMULATE", "-shared", "-fPIC", "-I", CSRC] + EXTRA_FLAGS + ["-o", so, cpp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed:\n" + r.stderr[-4000:])
    lib = C.CDLL(so)
    lib.pdl_k_info.argtypes = [C.POINTER(C.c_int)]
    lib.pdl_k_launch.argtypes = [C.POINTER(KernelParams), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pdl_k_launch.restype = C.c_int
    return lib, src


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def run_emu(desc, points, params, xi, mask, inv_n=None, backward=True, targets=None, n_ctas=2, rows_dev=None):
    """Returns dict(loss, sum_r2, residual, features, grad_params(flat), grad_xi)."""
    lib, _ = build_emu(desc)
    info = (C.c_int * 16)()
    lib.pdl_k_info(info)
    smem, nt, zscr, wpriv, cta_out, nprm, J, K, D, NW, nscal, mode = [int(info[i]) for i in range(12)]
    assert nprm == len(params), (nprm, len(params))
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n = pts.shape[0]
    prm = [np.ascontiguousarray(p, dtype=np.float32) for p in params]
    xi_a = np.ascontiguousarray(xi, dtype=np.float32) if K > 0 else np.zeros(1, np.float32)
    mask_a = np.ascontiguousarray(mask, dtype=np.uint8) if K > 0 else np.zeros(1, np.uint8)
    nwarps = n_ctas * NW
    P = KernelParams()
    P.points = ptr(pts); P.n_points = n
    for i, a in enumerate(prm):
        P.prm[i] = a.ctypes.data
    P.xi = ptr(xi_a); P.mask = ptr(mask_a)
    tg = np.ascontiguousarray(targets, dtype=np.float64) if targets is not None else None
    P.targets = ptr(tg)
    P.inv_n = float(inv_n if inv_n is not None else 1.0 / max(n, 1))
    resid = np.full(n, np.nan, np.float32); feats = np.full((J, n), np.nan, np.float32)
    P.residual = ptr(resid); P.features = ptr(feats)
    zs = np.full(max(1, nwarps * zscr) + 4, np.nan, np.float32)
    wp = np.full(max(1, nwarps * wpriv) + 4, np.nan, np.float32)
    co = np.full(max(1, n_ctas * cta_out), np.nan, np.float32)
    cs = np.zeros(n_ctas * 4, np.float64)
    # 16-byte alignment for float4 accesses
    def al(a):
        off = (-a.ctypes.data) % 16 // 4
        return a[off:]
    zs_a, wp_a = al(zs), al(wp)
    P.zscratch = ptr(zs_a); P.wpriv = ptr(wp_a); P.cta_out = ptr(co); P.cta_stats = ptr(cs)
    gp = np.full(nscal, np.nan, np.float32); gx = np.full(max(1, K), np.nan, np.float32); st = np.zeros(4, np.float64)
    cnt = np.array([int(rows_dev)], dtype=np.int64) if rows_dev is not None else None     # pdl_launch_args.n_points_dev
    P.n_points_dev = ptr(cnt)
    rc = lib.pdl_k_launch(C.byref(P), n_ctas, 1 if backward else 0, ptr(gp), ptr(gx), ptr(st), None)
    assert rc == 0
    return {"loss": st[0], "sum_r2": st[1], "sum_abs": st[2], "residual": resid, "features": feats,
            "grad_params": gp, "grad_xi": gx[:K], "J": J, "K": K}
