"""Deletes cached plan modules (pde_learn_b200/csrc/_plans) that no current configuration maps to.
Re-derives the cache key of runtime.cu (FNV-1a of generated source + flags + toolkit + kernel templates) for the build() menu, the
FP32-pipe plans and the kernel variants; anything else is stale (older templates, one-off A/B flags).  Extra flag sets to keep can be
given as arguments:  python scripts/plan_gc.py "-DPDL_TC_NWG=4" ...   (--dry to only list)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
from pde_learn_b200 import _lib as L

CSRC = os.path.join(ROOT, "pde_learn_b200", "csrc")
PLANS = os.path.join(CSRC, "_plans")
CUDART_VERSION = 12090
NVCC = os.environ.get("PDL_NVCC") or "/usr/local/cuda/bin/nvcc"


def fnv1a(b):
    h = 1469598103934665603
    for c in b:
        h ^= c
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def rd(name):
    return open(os.path.join(CSRC, name)).read()


def key_of(item, extra_flags=""):
    desc, _keep = L.make_desc(*item)
    src = L.plan_source(desc)
    engine1 = "jet_mlp_kernel_mma.cuh" in src
    hdr = rd("jet_mlp_kernel_mma.cuh" if engine1 else "jet_mlp_kernel.cuh") + rd("jet_common.cuh") + rd("kernel_params.h")
    if "#define PDL_FWD_TC 1" in src:
        hdr += rd("jet_mlp_fwd_tc.cuh")
    flags = "-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo " + extra_flags
    blob = src + "\n//" + flags + "\n//cuda " + str(CUDART_VERSION) + " " + NVCC + "\n" + hdr
    return "%016x" % fnv1a(blob.encode())


def current_keys(extra=()):
    keep = set()
    menu = g._plan_menu()
    for m in menu:
        keep.add(key_of(m))
    os.environ["PDL_ENGINE"] = "ffma"
    for m in menu:
        if m[1][1:-1] == [40] * 5 and m[3] == 0:
            keep.add(key_of(m))
    for m in g._variant_menu():
        keep.add(key_of(m))
    del os.environ["PDL_ENGINE"]
    from tests.fixtures_meta import KERNEL_VARIANTS
    for env, _tol in KERNEL_VARIANTS:
        os.environ.update(env)
        for m in g._variant_menu():
            keep.add(key_of(m))
        for k in env:
            del os.environ[k]
    for fl in extra:
        os.environ["PDL_NVCC_FLAGS"] = fl
        for m in menu:
            if m[1][1:-1] == [40] * 5 and m[3] == 0:
                keep.add(key_of(m, fl))
        del os.environ["PDL_NVCC_FLAGS"]
    return keep


def main():
    dry = "--dry" in sys.argv
    extra = [a for a in sys.argv[1:] if a != "--dry"]
    keep = current_keys(extra)
    have = sorted({f.split(".")[0][5:] for f in os.listdir(PLANS) if f.startswith("plan_")})
    stale = [h for h in have if h not in keep]
    print("plans on disk %d, current %d, missing %d, stale %d" % (len(have), len(keep), len(keep - set(have)), len(stale)))
    if not dry:
        for h in stale:
            for f in os.listdir(PLANS):
                if f.startswith("plan_" + h):
                    os.remove(os.path.join(PLANS, f))


if __name__ == "__main__":
    main()
