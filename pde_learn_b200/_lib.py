"""
ctypes binding of libpdelearn_b200.so (C ABI: include/pdelearn_b200.h).

The library is the product; this module only loads it and declares prototypes.  There is no Python or
CPU fallback: if the shared library is missing, importing the package raises with build instructions.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpdelearn_b200.so")

PDL_ACT_TANH, PDL_ACT_RATIONAL = 0, 1
PDL_MODE_COLLOCATION, PDL_MODE_DATA = 0, 1
PDL_MAX_PARAM_TENSORS = 40


class PlanDesc(C.Structure):
    _fields_ = [("d", C.c_int32), ("n_layers", C.c_int32), ("widths", C.POINTER(C.c_int32)),
                ("activation", C.c_int32), ("mode", C.c_int32), ("n_derivs", C.c_int32),
                ("encodings", C.POINTER(C.c_int32)), ("n_terms", C.c_int32), ("term_ptr", C.POINTER(C.c_int32)),
                ("term_deriv", C.POINTER(C.c_int32)), ("term_pow", C.POINTER(C.c_int32))]


class PlanInfo(C.Structure):
    _fields_ = [("d", C.c_int32), ("n_jets", C.c_int32), ("n_rhs_terms", C.c_int32), ("n_param_tensors", C.c_int32),
                ("n_param_scalars", C.c_int32), ("threads_per_cta", C.c_int32), ("smem_bytes", C.c_int32),
                ("warps_per_cta", C.c_int32), ("padded_width", C.c_int32), ("hidden_layers", C.c_int32),
                ("max_order", C.c_int32), ("mode", C.c_int32), ("flops_per_point", C.c_int64),
                ("jet_encodings", (C.c_int32 * 4) * 16)]


class LaunchArgs(C.Structure):
    _fields_ = [("points", C.c_void_p), ("n_points", C.c_int64), ("params", C.POINTER(C.c_void_p)),
                ("n_params", C.c_int32), ("xi", C.c_void_p), ("mask", C.c_void_p), ("targets", C.c_void_p),
                ("inv_n", C.c_double), ("backward", C.c_int32), ("residual", C.c_void_p), ("features", C.c_void_p),
                ("grad_params", C.c_void_p), ("grad_xi", C.c_void_p), ("stats", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("n_ctas", C.c_int32),
                ("stream", C.c_void_p), ("n_points_dev", C.c_void_p)]


# every symbol include/pdelearn_b200.h declares, with its prototype
PROTOTYPES = {
    "pdl_last_error": (C.c_char_p, []),
    "pdl_abi_version": (C.c_int, []),
    "pdl_plan_source": (C.c_int, [C.POINTER(PlanDesc), C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "pdl_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]),
    "pdl_plan_destroy": (None, [C.c_void_p]),
    "pdl_plan_info_get": (C.c_int, [C.c_void_p, C.POINTER(PlanInfo)]),
    "pdl_plan_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32]),
    "pdl_device_sm_count": (C.c_int, [C.POINTER(C.c_int32)]),
    "pdl_loss_launch": (C.c_int, [C.c_void_p, C.POINTER(LaunchArgs)]),
    "pdl_lp_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdl_l2_squared_loss": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "pdl_sample_points": (C.c_int, [C.POINTER(C.c_double), C.c_int32, C.c_int64, C.c_uint64, C.c_uint64,
                                    C.c_void_p, C.c_void_p]),
    "pdl_sample_points_dev": (C.c_int, [C.POINTER(C.c_double), C.c_int32, C.c_int64, C.c_void_p, C.c_uint64,
                                        C.c_void_p, C.c_void_p]),
    "pdl_abs_residual_moments": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pdl_philox4x32_10_raw": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pdl_select_targeted": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdl_targeted_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdl_targeted_update_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdl_select_targeted_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdl_adam_step_multi": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_int64), C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.c_void_p, C.c_void_p]),
    "pdl_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p]),
    "pdl_axpby3": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_double,
                             C.c_void_p]),
    "pdl_closure_tail": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "pdl_release_l2_setaside": (C.c_int, []),
    "pdl_measure_fp32_peak": (C.c_int, [C.POINTER(C.c_double), C.c_void_p]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """The loaded shared library (loads on first use; raises loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                "libpdelearn_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C pde_learn_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)            # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.pdl_abi_version() != 3:
            raise NativeLibraryError("libpdelearn_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise NativeLibraryError(lib().pdl_last_error().decode("utf-8", "replace"))


def make_desc(d, widths, activation, mode, encodings, terms):
    """
    Build a PlanDesc.  encodings: list of tuples (len 2..4).  terms: list (LHS first) of lists of
    (encoding index, power).  Returns (desc, keepalive) -- keep `keepalive` referenced while desc is used.
    """
    w = (C.c_int32 * len(widths))(*[int(v) for v in widths])
    enc_flat = []
    for e in encodings:
        e = [int(v) for v in e]
        enc_flat += e + [0] * (4 - len(e))
    enc = (C.c_int32 * max(1, len(enc_flat)))(*enc_flat)
    ptr, der, pw = [0], [], []
    for t in terms:
        for (di, p) in t:
            der.append(int(di)); pw.append(int(p))
        ptr.append(len(der))
    tptr = (C.c_int32 * len(ptr))(*ptr)
    tder = (C.c_int32 * max(1, len(der)))(*der)
    tpow = (C.c_int32 * max(1, len(pw)))(*pw)
    desc = PlanDesc(int(d), len(widths) - 1, w, int(activation), int(mode), len(encodings), enc, len(terms),
                    tptr, tder, tpow)
    return desc, (w, enc, tptr, tder, tpow)


def plan_source(desc):
    need = C.c_size_t(0)
    check(lib().pdl_plan_source(C.byref(desc), None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    check(lib().pdl_plan_source(C.byref(desc), buf, need.value, C.byref(need)))
    return buf.value.decode()
