"""
Memory-safety check of the kernel source without a GPU (compute-sanitizer is closed on this pool): the emulated kernel
(-DPDL_EMULATE) is rebuilt with AddressSanitizer + UBSan and run in a subprocess with libasan preloaded.  Every shared,
scratch and global access of the kernel template goes through ordinary C++ pointers there, so an out-of-bounds index in
the tiling code aborts the subprocess.  Buffers are allocated at their exact sizes (tests/emu.py).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from tests import emu
emu.EXTRA_FLAGS = ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"]
emu.EMU_DIR = emu.EMU_DIR + "_asan"
from tests.helpers import GoldenCase, rel_max
from tests.test_emulation_parity import desc_for
from pde_learn_b200 import _lib as L
for name, n, ctas in (("ks_rational", 21, 2), ("rd2d_tanh", 9, 1), ("heat3d_tanh", 17, 3), ("wave_mixed_tanh", 33, 2)):
    g = GoldenCase(name)
    desc, keep = desc_for(g)
    out = emu.run_emu(desc, g.points.numpy()[:n], [p.numpy() for p in g.params], g.xi.numpy(), np.array(g.mask, dtype=np.uint8), n_ctas=ctas)
    assert rel_max(out["residual"], g.z["residual"][:n]) < 1e-5
    out = emu.run_emu(desc, g.points.numpy()[:n], [p.numpy() for p in g.params], g.xi.numpy(), np.array(g.mask, dtype=np.uint8), n_ctas=ctas, backward=False)
    d2, k2 = desc_for(g, L.PDL_MODE_DATA)
    emu.run_emu(d2, g.x_data.numpy()[:n], [p.numpy() for p in g.params], None, None, targets=g.y_data.numpy()[:n], n_ctas=ctas)
print("ASAN_OK")
"""


def test_emulated_kernel_is_clean_under_asan_ubsan():
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan not available")
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1:halt_on_error=1")
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "ASAN_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
