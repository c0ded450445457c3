"""
Golden vectors for the data-set preparation step (reference: Data/From_MATLAB.py:37-263, Data/Create_Data_Set.py:5-56).

Builds two small synthetic MATLAB files (1-D: usol[Nx,Nt]; 2-D: usol[Nt,Nx,Ny], the layouts of the reference's Matlab/Data files),
runs the UNMODIFIED reference functions on them in a scratch directory (matplotlib is absent here and is stubbed; plotting is
switched off) with numpy.random.seed(11), and stores what they wrote.  Run in the build container only:
    python tests/golden/make_golden_dataset.py
"""
import os
import shutil
import sys
import tempfile
import types

import numpy
import scipy.io

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Data"


def main():
    rng = numpy.random.default_rng(5)
    t = numpy.linspace(0.0, 2.0, 18).reshape(1, -1); x = numpy.linspace(-3.0, 3.0, 24).reshape(1, -1)
    u1 = numpy.sin(x.reshape(-1, 1)) * numpy.exp(-0.3 * t.reshape(1, -1)) + 0.01 * rng.standard_normal((24, 18))
    scipy.io.savemat(os.path.join(HERE, "synthetic_1d.mat"), {"t": t, "x": x, "usol": u1})
    t2 = numpy.linspace(0.0, 1.0, 6).reshape(1, -1); x2 = numpy.linspace(-1.0, 1.0, 8).reshape(1, -1); y2 = numpy.linspace(0.0, 2.0, 7).reshape(1, -1)
    u2 = (numpy.exp(-t2.reshape(-1, 1, 1)) * numpy.cos(x2.reshape(1, -1, 1)) * numpy.sin(y2.reshape(1, 1, -1))
          + 0.01 * rng.standard_normal((6, 8, 7)))
    scipy.io.savemat(os.path.join(HERE, "synthetic_2d.mat"), {"t": t2, "x": x2, "y": y2, "usol": u2})

    stub = types.ModuleType("matplotlib"); stub.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules["matplotlib"] = stub; sys.modules["matplotlib.pyplot"] = stub.pyplot
    sys.path.insert(0, REF)
    import From_MATLAB as F
    F.Make_Plot = False
    work = tempfile.mkdtemp()
    try:
        os.makedirs(os.path.join(work, "MATLAB", "Data")); os.makedirs(os.path.join(work, "Data", "DataSets"))
        for n in ("synthetic_1d", "synthetic_2d"):
            shutil.copy(os.path.join(HERE, n + ".mat"), os.path.join(work, "MATLAB", "Data", n + ".mat"))
        os.chdir(os.path.join(work, "Data"))
        numpy.random.seed(11)
        F.From_MATLAB_1D("synthetic_1d", 0.25, 120, 40)
        numpy.random.seed(11)
        F.From_MATLAB_2D("synthetic_2d", 0.25, 90, 30)
        for n, out in (("synthetic_1d_N25_P120", "dataset_1d.npz"), ("synthetic_2d_N25_P90", "dataset_2d.npz")):
            shutil.copy(os.path.join(work, "Data", "DataSets", n + ".npz"), os.path.join(HERE, out))
            z = numpy.load(os.path.join(HERE, out))
            print(out, {k: (z[k].shape, str(z[k].dtype)) for k in z.files})
    finally:
        os.chdir(HERE)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
